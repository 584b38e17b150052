// Tensor-core path of the margin head for sm_100a: tcgen05.mma (cta_group::2, one SM pair per tile) with TMEM
// accumulators, TMA-fed 128B-swizzled shared-memory stages, mbarrier pipelines, warp-specialised persistent CTAs.
//
// One kernel skeleton, five modes (template parameter):
//   FWD    S = x^ w^T tile -> clamp/margin/scale -> sum-exp + rank count per row.  Nothing of size B x C is
//          written (replaces criterion.py:267-301 & siblings + nn.CrossEntropyLoss, model_utils.py:179,
//          + accuracy, metrics.py:3-16).
//   FWDS   FWD that also stashes E' = exp2(z log2e - ref_i) * du/dcos as bf16 (target column and padding = 0), so a
//          training step needs no logit recompute: G_ij = rho_i E'_ij with rho_i known once the row's lse is.
//   BWD_G  recompute the S tile -> G = (P - Y) * dz/dcos as bf16 (the recompute backward).
//          Both B x C layouts are class-tiled [C_pad/128][B_pad][128].
//   DX     dx^ partials = G . w^      (A = G K-major,  B = w^ MN-major, split over classes; one 128 x 512
//          accumulator = all 512 TMEM columns per CTA, so every G byte is read by exactly one CTA)
//          In stash mode the otherwise idle epilogue warps also read every A tile after its MMAs retired and
//          accumulate r_j = sum_i G_ij cos_ij (= w^_j . dw^_j, the projection of the normalise-backward of W),
//          recovering cos_ij from the stashed exponential itself.
//   DW     dw^ = G^T . x^  (A = G MN-major, B = x^ MN-major), 256 classes x 256 d per pair tile; the epilogue writes
//          dW_j = g (dw^_j - w^_j r_j) / |w_j| straight into the parameter layout (or raw dw^ on request).
//
// CTA = 384 threads: warp 0 TMA producer, warp 1 MMA issuer (leader CTA), warp 2 TMEM allocator, warps 4-11
// epilogue (thread = one accumulator row x one column half).
#include "common.cuh"
#include <cuda.h>
#include <cstdlib>
#include <mutex>
#include <algorithm>

namespace {

constexpr int BM = 128, BN = 256, BK = 64, MAX_STAGES = 6;
constexpr int BMT = 2 * BM;                       // rows of a pair tile
constexpr int A_STAGE_BYTES = BM * BK * 2;        // 16 KB
constexpr int NUM_EPI_WARPS = 8;
constexpr int EPI_WARP0 = 4;
constexpr int NUM_THREADS = 32 * (EPI_WARP0 + NUM_EPI_WARPS);
constexpr uint32_t TMEM_COLS = 512;
constexpr int PROG_AHEAD = 16;                    // merged dx+dW kernel: max lead of one role over the other, in 256-class tiles
constexpr int DX_SYNC_KB = 16;                    // DX lockstep: k-blocks between two rendezvous of a split's CTAs
constexpr int STG_WARP_BYTES = 32 * 128;          // output staging per epilogue warp: 32 rows x 128 B, XOR-swizzled

enum { MODE_FWD = 0, MODE_FWDS = 1, MODE_BWD_G = 2, MODE_DX = 3, MODE_DW = 4 };
// A-stationary (FWD / FWDS / BWD_G): the pair's x^ tile (128 rows x 512 per CTA = 128 KB) stays resident in shared
// memory while the pair sweeps class tiles, so only w^ streams (16 KB per CTA per k-block): L2 -> SM traffic per
// tile halves (64 -> 32 B/clk/SM; the non-stationary kernels sat on the ~6.3 KB/clk chip-wide L2 delivery cap).
constexpr bool mode_astat(int mode) { return mode == MODE_FWD || mode == MODE_FWDS || mode == MODE_BWD_G; }
constexpr bool mode_staged(int mode) { return mode == MODE_FWDS || mode == MODE_BWD_G || mode == MODE_DW; }
constexpr int A_RESIDENT_BYTES = BM * MH_D * 2;                                   // 128 KB
constexpr int mode_bn(int mode) { return mode == MODE_DX ? 512 : 256; }          // accumulator columns per tile
constexpr int mode_nbuf(int mode) { return mode_bn(mode) == 512 ? 1 : 2; }        // TMEM accumulators in flight
constexpr int mode_stage_bytes(int mode) { return (mode_astat(mode) ? 0 : A_STAGE_BYTES) + (mode_bn(mode) / 2) * BK * 2; }
constexpr int mode_stages(int mode) { return mode == MODE_FWD ? 6 : 4; }
// DW: the epilogue's w^ tile (128 classes x 256 d, bf16) is TMA-loaded into shared memory once per tile
constexpr int W_TILE_BYTES = BM * BN * 2;                                         // 64 KB
constexpr int mode_smem_bytes(int mode) {
  return (mode_astat(mode) ? A_RESIDENT_BYTES : 0) + mode_stages(mode) * mode_stage_bytes(mode) + 1024 /*align slack*/ +
         256 /*barriers*/ + (mode_staged(mode) ? NUM_EPI_WARPS * STG_WARP_BYTES : 0) +
         (mode == MODE_DW ? W_TILE_BYTES : 0);
}
static_assert(mode_smem_bytes(MODE_FWD) <= 232448 && mode_smem_bytes(MODE_FWDS) <= 232448 &&
              mode_smem_bytes(MODE_BWD_G) <= 232448 && mode_smem_bytes(MODE_DX) <= 232448 &&
              mode_smem_bytes(MODE_DW) <= 232448, "smem budget");

struct TcArgs {
  CUtensorMap tmW;             // DW fused: w^ [C_pad, 512], box [64 d][128 classes] (epilogue operand)
  int m_tiles, n_tiles, n_split, k_blocks_total, k_blocks_per_split;
  int64_t total_tiles;
  int sG, sE, n_fixed;         // A-stationary schedule (see StatIter)
  MhParams p;
  int64_t B, C, B_pad, C_pad;
  const float* rowp;
  int64_t ldp;
  const int32_t* label_local;
  const float* state;
  const float* lse2;
  float* stats_tiles;
  __nv_bfloat16* G;            // BWD_G: G out; FWDS: stash out
  float* out;
  int64_t out_split_stride;
  int fixref;                  // FWD: fixed softmax reference (see mh_tc_fixref_ok); always 1 for FWDS
  float umax;                  // upper bound of u = z / scale over non-target columns (fixref)
  float* rsum;                 // r_j: BWD_G accumulates [C_pad] (atomics); DX (stash) stores one partial per 128-row block
                               // ([B_pad/128][C_pad], plain stores: reproducible); DW fused sums rsum_parts partials
  int rsum_parts;              // DW fused: number of partial planes of rsum (1 for the BWD_G sums)
  float* rpart;                // DW self-projection: [4][C_pad] partial dots w^_j . dw^_j (d half x epilogue column half)
  int* rflag;                  // DW self-projection: [C_pad/128] arrival counters, zeroed before the launch
  int* dx_sync;                // DX lockstep: [n_split] arrival counters (NULL: off), zeroed before the launch
  int dx_chunk;                // DX: k-blocks per interleaved chunk (0: contiguous splits)
  int* prog;                   // merged dx+dW kernel: prog[0] = dx front, prog[1] = dW front (256-class tiles); NULL: off
  int prog_ahead;              // merged dx+dW kernel: max lead of one role over the other, in 256-class tiles
  const int* w_ready;          // merged prologue+forward kernel: per 256-class tile, #rows of w^ written so far (NULL: off)
  int* fwd_front;              // merged prologue+forward kernel: class tile the forward leader pair is loading
  const float* rho;            // DX side pass: rho_i of the stash rows (NULL: no side pass)
  float side_kappa, side_inv_s2;   // DX side pass: cos = log2(E') * inv_s2 + kappa
  int side_mv;                 // DX side pass, MV-Softmax: invert the hard-negative re-weighting u = a*c + b as well
  float side_ha, side_hb;      //   (a, b) = (mv_weight, mv_weight - 1)
  const __nv_bfloat16* w_hat;  // DW fused
  const float* inv_norm;       // DW fused
  const float* gscal;          // DW fused
  int raw_dw;                  // DW: write raw dw^ [C_pad, 512] instead of the projected dW
  int layout;                  // DW fused: parameter layout of dW
  int64_t ld;                  // DW fused: row pitch of dW
  const int* gate;             // guarded stash (mh_step_*): the launch does nothing unless (*gate != 0) == (gate_on != 0); NULL: always runs
  int gate_on;
};

// Guarded-stash fallback kernels are launched unconditionally and decide on the device (no host sync): every thread of
// every CTA reads the same flag, which no kernel in flight writes, so the exit is uniform across the cluster.
__device__ __forceinline__ bool gate_closed(const int* gate, int gate_on) {
  return gate != nullptr && ((*reinterpret_cast<const volatile int*>(gate) != 0) != (gate_on != 0));
}

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
// Long waits (the dx epilogue warps wait ~1 ms for a whole split-K accumulation): back off with nanosleep instead of
// re-issuing try_wait at full rate - fewer wasted issue slots next to the MMA / producer warps, same wake-up within 1 us.
__device__ __forceinline__ void mbar_wait_long(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  unsigned ns = 64;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(ns);
    ns = ns < 1024 ? ns * 2 : 1024;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// TMA load into this CTA's shared memory, completing on this CTA's mbarrier
__device__ __forceinline__ void tma_load_2d_local(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float lg2(float x) {      // log2, no denormal fix-up: lg2(0) = -inf
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// named barrier among the epilogue warps only (barrier 0 is __syncthreads)
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * NUM_EPI_WARPS) : "memory"); }
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu(int* p, int v) {
  asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_relaxed_gpu_add(int* p, int v) {
  asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- cta_group::2 (SM pair) variants -----------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(rank)
      : "memory");
}
// TMA load issued by either CTA of the pair; the transaction bytes land on the LEADER CTA's mbarrier
// (peer bit 24 of the shared::cluster address cleared).
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
// L2 eviction-priority policies for the TMA loads (cache_hint operand) and streaming stores - EXPERIMENT, compiled in
// with -DMH_L2_HINTS only.  Idea: the B x C stream (stash / G) is read once per consumer and pushes the operand that IS
// re-read by neighbouring pairs (w^, x^) out of L2.  Measured (profiles/r2_ab_merged_hints.txt, ncu): evict-first on the
// stash + evict-last on w^ RAISED the dx kernel's DRAM reads from 7.65 to 9.17 GB and made the merged backward slower
// than the two separate kernels; without hints the merged kernel is 0.4 ms faster.  Off by default.
#ifdef MH_L2_HINTS
#define MH_STCS(ptr, val) __stcs((ptr), (val))
#else
#define MH_STCS(ptr, val) (*(ptr) = (val))
#endif
#ifdef MH_L2_HINTS
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
#else
__device__ __forceinline__ uint64_t l2_policy_evict_first() { return 0; }
__device__ __forceinline__ uint64_t l2_policy_evict_last() { return 0; }
#endif
// as tma_load_2d_2sm, with an L2 cache policy (0 = none)
__device__ __forceinline__ void tma_load_2d_2sm_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                     uint64_t policy) {
#ifdef MH_L2_HINTS
  if (policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
    return;
  }
#endif
  tma_load_2d_2sm(dst, map, bar, c0, c1);
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// commit of the pair's MMAs: arrives on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
      ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// ---- UMMA descriptors -------------------------------------------------------------------------------
// Shared-memory matrix descriptor (sm_100): start addr [0,14) >>4, LBO [16,30) >>4, SBO [32,46) >>4,
// version=1 at [46,48), layout type [61,64) (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// K-major tile [rows][64 bf16] (128 B per row, 8-row swizzle atoms of 1024 B): SBO = 1024, LBO unused.
// One UMMA (K=16) advances the start address by 32 B inside the swizzle atom.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_base, int k16) {
  return make_desc(tile_base + k16 * 32, 16, 1024);
}
// MN-major operand built from [64 K-rows][64 MN elems] boxes (8 KB each, box b covers MN 64b..64b+63):
// LBO = 8192 (next 64-wide MN block), SBO = 1024 (next group of 8 K rows); K=16 -> +2048 B.
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_base, int k16) {
  return make_desc(tile_base + k16 * 2048, 8192, 1024);
}
// Instruction descriptor, kind::f16: D=F32 (bit 4), A=B=BF16 (bits 7,10), majors (15,16), N>>3 (17..22), M>>4 (24..28).
__host__ __device__ constexpr uint32_t make_idesc(int a_mn, int b_mn, int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Work {
  int m0, n0;        // output tile origin (rows of D, cols of D)
  int kb0, kb1;      // k-block range
  int kb_chunk, kb_jump;   // DX interleaved split (merged dx+dW kernel): kb_chunk k-blocks every kb_jump; 0 = contiguous
  int split;         // DX split index
  int n_tile;        // FWD: class-tile index
  int m_tile;        // A-stationary modes: row-tile index (the resident x^ tile)
};

// A-stationary tile schedule (FWD / FWDS / BWD_G).  `units` CTA pairs, m_tiles <= units row tiles of 256 rows:
//   * sG = units / m_tiles pairs are bound to each row tile m; pair q of the group takes class tiles q, q+sG, ... of
//     [0, n_fixed).  The groups of all row tiles sweep the classes in lockstep, so a w^ tile fetched from HBM by one
//     row tile is an L2 hit for the others.
//   * the sE = units - sG*m_tiles left-over pairs share the class tiles [n_fixed, n_tiles) of ALL row tiles, each taking
//     a contiguous chunk in (m-major, n-minor) order (they reload x^ when m changes); n_fixed balances both kinds.
struct StatIter {
  int fixed, m, n, step, n_end, u, u_end, n_ext, n_fixed;
  __host__ __device__ __forceinline__ void init(const TcArgs& a, int pid) {
    const int nfix_pairs = a.sG * a.m_tiles;
    fixed = pid < nfix_pairs;
    m = n = step = n_end = u = u_end = 0;
    n_fixed = a.n_fixed;
    n_ext = a.n_tiles - a.n_fixed;
    if (fixed) {
      m = pid % a.m_tiles; n = pid / a.m_tiles; step = a.sG; n_end = a.n_fixed;
    } else if (a.sE > 0 && n_ext > 0) {
      const int e = pid - nfix_pairs;
      const int64_t U = (int64_t)n_ext * a.m_tiles;
      u = (int)(U * e / a.sE);
      u_end = (int)(U * (e + 1) / a.sE);
    }
  }
  __host__ __device__ __forceinline__ bool next(int& m_out, int& n_out) {
    if (fixed) {
      if (n >= n_end) return false;
      m_out = m; n_out = n; n += step;
      return true;
    }
    if (u >= u_end) return false;
    m_out = u / n_ext; n_out = n_fixed + u % n_ext; ++u;
    return true;
  }
};

// next k-block of a Work: contiguous [kb0, kb1), or chunks of kb_chunk k-blocks every kb_jump (kb_jump % kb_chunk == 0)
__device__ __forceinline__ int kb_next(const Work& w, int kb) {
  ++kb;
  if (w.kb_chunk && (kb - w.kb0) % w.kb_chunk == 0) kb += w.kb_jump - w.kb_chunk;
  return kb;
}

// The tile sequence of one CTA pair; the producer, MMA and epilogue roles all walk the same sequence.
template <int MODE>
struct TileLoop {
  StatIter si;
  int64_t t, npid;
  __device__ __forceinline__ void init(const TcArgs& a, int64_t pid, int64_t npid_) {
    t = pid; npid = npid_;
    if (mode_astat(MODE)) si.init(a, (int)pid);
  }
  __device__ __forceinline__ bool next(const TcArgs& a, int rank, Work& w) {
    w.split = 0; w.n_tile = 0; w.m_tile = 0; w.kb_chunk = 0; w.kb_jump = 0;
    if (mode_astat(MODE)) {
      int m, n;
      if (!si.next(m, n)) return false;
      w.m0 = m * BMT + rank * BM; w.n0 = n * BN; w.kb0 = 0; w.kb1 = MH_D / BK;
      w.n_tile = n; w.m_tile = m;
      return true;
    }
    if (t >= a.total_tiles) return false;
    if (MODE == MODE_DX) {
      w.split = (int)(t / a.m_tiles);
      w.m0 = (int)(t % a.m_tiles) * BMT + rank * BM;
      w.n0 = 0;
      if (a.dx_chunk) {          // interleaved: split s takes chunks s, s + n_split, ... so that all pairs walk the classes together
        w.kb_chunk = a.dx_chunk; w.kb_jump = a.dx_chunk * a.n_split;
        w.kb0 = w.split * a.dx_chunk; w.kb1 = a.k_blocks_total;
      } else {
        w.kb0 = w.split * a.k_blocks_per_split;
        w.kb1 = min(a.k_blocks_total, w.kb0 + a.k_blocks_per_split);
      }
    } else {  // DW: 256 classes x one 256-wide half of d
      w.m0 = (int)(t >> 1) * BMT + rank * BM;
      w.n0 = (int)(t & 1) * BN;
      w.kb0 = 0; w.kb1 = a.k_blocks_total;
    }
    t += npid;
    return true;
  }
};

// ---- epilogue helpers -----------------------------------------------------------------------------
// Family variants of the B x C element transform (compile-time, so the hot loop carries no dead work):
//   V_PLAIN  no clamp                       (ArcFace, criterion.py:267-301)
//   V_CLAMP  clamp only                     (CosFace, AdaFace, ElasticFace, MagFace)
//   V_SPHERE clamp + sum e*u statistic      (SphereFace: logits scale with |x|, criterion.py:105)
//   V_MV     clamp + c>thr ? w*c+w-1 : c    (MV-Softmax, criterion.py:433-435)
//   V_CURR   clamp + c>thr ? c*(t+c) : c    (CurricularFace, criterion.py:559,575)
enum { V_PLAIN = 0, V_CLAMP = 1, V_SPHERE = 2, V_MV = 3, V_CURR = 4, V_NONE = 5 };

struct RowCtx {
  float scale, scale2, thr, t, zt2, dzt, lse2;
  float nref2;       // fixref: -(softmax reference) in log2 units = 102 - scale2 * umax
  float ntbig;       // fixref: -t * 2^60 (rank count through FFMA.SAT)
  int tcol;          // tile-local target column, or -1
  bool valid;        // row < B
};

struct FwdAcc {
  float m, l, ez;
  float cntf;
  int cnt;
};

constexpr float CNT_BIG = 1152921504606846976.f;   // 2^60

template <int V, bool CL = true>
__device__ __forceinline__ float elem_u(float raw, float lo, float hi, float thr, float ha, float hb, float& c_out) {
  float c = raw;
  if (V != V_PLAIN && CL) c = fminf(fmaxf(raw, lo), hi);
  c_out = c;
  if (V == V_MV) return (c > thr) ? fmaf(ha, c, hb) : c;
  if (V == V_CURR) return (c > thr) ? c * (ha + c) : c;
  return c;
}
template <int V, bool CL = true>
__device__ __forceinline__ float elem_du(float raw, float c, float thr, float ha) {
  float du = 1.f;
  if (V == V_MV) du = (c > thr) ? ha : 1.f;
  if (V == V_CURR) du = (c > thr) ? (ha + 2.f * c) : 1.f;
  if (V != V_PLAIN && CL && c != raw) du = 0.f;      // clamp passes gradient only inside [lo, hi]
  return du;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// One 32-column chunk of the forward, general form: online max / sum-exp (log2 domain), rank count, optional sum e*u.
template <int V>
__device__ __forceinline__ void fwd_chunk(uint32_t (&v)[32], int col0, int nvalid, const RowCtx& rc, float lo,
                                          float hi, float ha, float hb, FwdAcc& acc) {
  const bool slow = (rc.tcol >= col0 && rc.tcol < col0 + 32) || (col0 + 32 > nvalid);
  float umax = -INFINITY;
  if (!slow) {
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      float c;
      const float u = elem_u<V>(__uint_as_float(v[k]), lo, hi, rc.thr, ha, hb, c);
      if (c > rc.t) acc.cnt += 1;
      umax = fmaxf(umax, u);
      v[k] = __float_as_uint(u);
    }
    const float zmax = umax * rc.scale2;                 // scale2 >= 0, so max z = scale2 * max u
    if (zmax > acc.m) {
      const float rs = ex2(acc.m - zmax);
      acc.l *= rs; acc.ez *= rs; acc.m = zmax;
    }
    const float negm = -acc.m;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float u = __uint_as_float(v[k]);
      const float e = ex2(fmaf(u, rc.scale2, negm));
      acc.l += e;
      if (V == V_SPHERE) acc.ez = fmaf(e, u, acc.ez);
    }
  } else {
    // rare path: the chunk holds this row's target column and/or padded classes
    float z2[32];
    float zmax = -INFINITY;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      float c;
      const float u = elem_u<V>(__uint_as_float(v[k]), lo, hi, rc.thr, ha, hb, c);
      float z = u * rc.scale2;
      const int col = col0 + k;
      if (col == rc.tcol) z = rc.zt2;
      else if (col < nvalid && c > rc.t) acc.cnt += 1;
      if (col >= nvalid) z = -INFINITY;
      z2[k] = z;
      zmax = fmaxf(zmax, z);
    }
    if (zmax > acc.m) {
      const float rs = ex2(acc.m - zmax);
      acc.l *= rs; acc.ez *= rs; acc.m = zmax;
    }
    if (acc.m > -INFINITY) {
      const float inv_s2 = (rc.scale2 != 0.f) ? 1.f / rc.scale2 : 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const float e = ex2(z2[k] - acc.m);
        acc.l += e;
        if (V == V_SPHERE) acc.ez = fmaf(e, fmaxf(z2[k], -1e30f) * inv_s2, acc.ez);   // u = z/scale; 0 * -inf guard
      }
    }
  }
}

// One 32-column chunk of the forward with a FIXED softmax reference (no running max, no rescale):
// e = exp2(u*scale2 - ref2_i) with ref2_i = scale2_i*umax - 102, valid when every possible non-target term is a normal
// fp32/bf16 number (mh_tc_fixref_ok).  5 issue slots per element: FFMA, MUFU.EX2, FADD (sum), FFMA.SAT, FADD (rank
// count: sat((c - t) * 2^60) is exactly 1 for c > t, else 0).  STASH: E' = e * du/dcos -> 16 packed bf16 pairs.
template <int V, bool STASH>
__device__ __forceinline__ void fwd_chunk_fix(const uint32_t (&v)[32], int col0, int nvalid, const RowCtx& rc, float lo,
                                              float hi, float ha, float hb, FwdAcc& acc, uint32_t (&pk)[16]) {
  bool slow = (rc.tcol >= col0 && rc.tcol < col0 + 32) || (col0 + 32 > nvalid) || (STASH && !rc.valid);
  if (V != V_PLAIN && !slow) {
    // Clamped families: a cosine outside [lo, hi] needs near-duplicate vectors, so instead of clamping every element (two
    // FMNMX, plus a compare + select for the clamp's gradient mask) the chunk takes one max-|raw| pass (one FMNMX with a
    // free |.| modifier per element) and goes to the rare path, which clamps, if anything is out of range.
    float am = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) am = fmaxf(am, fabsf(__uint_as_float(v[k])));
    slow = am > fminf(-lo, hi);
  }
  if (!slow) {
    float l2[2] = {0.f, 0.f}, c2[2] = {0.f, 0.f}, z2[2] = {0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 32; k += 2) {
      float es[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float raw = __uint_as_float(v[k + h]);
        float c;
        const float u = elem_u<V, false>(raw, lo, hi, rc.thr, ha, hb, c);
        c2[h] += __saturatef(fmaf(c, CNT_BIG, rc.ntbig));
        const float e = ex2(fmaf(u, rc.scale2, rc.nref2));
        l2[h] += e;
        if (V == V_SPHERE) z2[h] = fmaf(e, u, z2[h]);
        es[h] = STASH ? e * elem_du<V, false>(raw, c, rc.thr, ha) : e;
      }
      if (STASH) pk[k >> 1] = pack_bf16(es[0], es[1]);
    }
    acc.l += l2[0] + l2[1];
    acc.cntf += c2[0] + c2[1];
    if (V == V_SPHERE) acc.ez += z2[0] + z2[1];
  } else {
#pragma unroll
    for (int k = 0; k < 32; k += 2) {
      float es[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float raw = __uint_as_float(v[k + h]);
        float c;
        const float u = elem_u<V>(raw, lo, hi, rc.thr, ha, hb, c);
        const int col = col0 + k + h;
        float e = ex2(fmaf(u, rc.scale2, rc.nref2));
        float sv = e * elem_du<V>(raw, c, rc.thr, ha);
        float ue = u;
        if (col == rc.tcol) { e = ex2(rc.zt2 + rc.nref2); sv = 0.f; ue = (rc.scale2 != 0.f) ? rc.zt2 / rc.scale2 : 0.f; }
        else if (col < nvalid && c > rc.t) acc.cnt += 1;
        if (col >= nvalid) { e = 0.f; sv = 0.f; }
        if (!rc.valid) sv = 0.f;
        acc.l += e;
        if (V == V_SPHERE) acc.ez = fmaf(e, ue, acc.ez);
        es[h] = sv;
      }
      if (STASH) pk[k >> 1] = pack_bf16(es[0], es[1]);
    }
  }
}

// Column sums over the 32 lanes of a warp of a 32-register array: on return lane L holds sum_lanes x[L].
// Transposed butterfly: 31 shuffles, no shared memory.
__device__ __forceinline__ float warp_colsum32(float (&x)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      const float send = upper ? x[j] : x[j + s];
      const float keep = upper ? x[j + s] : x[j];
      x[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return x[0];
}

// One 32-column chunk of the backward-G kernel: G = (P - Y) * dz/dcos -> 16 packed bf16 pairs, and
// q = G * cos_raw left in qv[] for the column sums r_j = sum_i G_ij cos_ij (= w^_j . dw^_j).
template <int V>
__device__ __forceinline__ void bwd_chunk(const uint32_t (&v)[32], int col0, int nvalid, const RowCtx& rc, float lo,
                                          float hi, float ha, float hb, uint32_t (&pk)[16], float (&qv)[32]) {
  const bool slow = (rc.tcol >= col0 && rc.tcol < col0 + 32) || (col0 + 32 > nvalid) || !rc.valid;
  const float negl = -rc.lse2;
  if (!slow) {
#pragma unroll
    for (int k = 0; k < 32; k += 2) {
      float g2[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float raw = __uint_as_float(v[k + h]);
        float c;
        const float u = elem_u<V>(raw, lo, hi, rc.thr, ha, hb, c);
        const float pr = ex2(fmaf(u, rc.scale2, negl));
        g2[h] = pr * rc.scale * elem_du<V>(raw, c, rc.thr, ha);
        qv[k + h] = g2[h] * raw;
      }
      pk[k >> 1] = pack_bf16(g2[0], g2[1]);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 32; k += 2) {
      float g2[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float raw = __uint_as_float(v[k + h]);
        float c;
        const float u = elem_u<V>(raw, lo, hi, rc.thr, ha, hb, c);
        float z = u * rc.scale2, dzdc = rc.scale * elem_du<V>(raw, c, rc.thr, ha), yv = 0.f;
        const int col = col0 + k + h;
        if (col == rc.tcol) { z = rc.zt2; dzdc = rc.dzt; yv = 1.f; }
        float g = (ex2(z + negl) - yv) * dzdc;
        if (col >= nvalid || !rc.valid) g = 0.f;
        g2[h] = g;
        qv[k + h] = g * raw;
      }
      pk[k >> 1] = pack_bf16(g2[0], g2[1]);
    }
  }
}

// Stage one 64 B half-row (16 packed bf16 pairs) of a [32 rows][128 B] warp staging tile; 16 B pieces XOR-swizzled
// with the row so that both the row-wise writes and the 4-rows-per-instruction reads are bank-conflict free.
__device__ __forceinline__ void stage_half_row(uint8_t* stg, int lane, int half64, const uint32_t (&pk)[16]) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
    *reinterpret_cast<uint4*>(stg + lane * 128 + (((half64 * 4 + k) ^ (lane & 7)) * 16)) =
        make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
}
// Flush a staged [32 rows][128 B] tile: 8 lanes per row (full 128 B lines), 4 rows per instruction.
// dst = address of row 0 / byte 0 of the tile in global memory, pitch in bytes; rows >= rows_ok are skipped.
__device__ __forceinline__ void flush_tile(const uint8_t* stg, int lane, uint8_t* dst, int64_t pitch_bytes, int rows_ok) {
#pragma unroll
  for (int i2 = 0; i2 < 8; ++i2) {
    const int rr = 4 * i2 + (lane >> 3);
    const uint4 val = *reinterpret_cast<const uint4*>(stg + rr * 128 + (((lane & 7) ^ (rr & 7)) * 16));
    // written once, not re-read by this kernel: streaming store (evict-first in L2), keeps the operands resident
    if (rr < rows_ok) MH_STCS(reinterpret_cast<uint4*>(dst + (int64_t)rr * pitch_bytes + (lane & 7) * 16), val);
  }
}

// Walk the NCHUNK 32-column chunks of this thread's accumulator row: the TMEM load of chunk c+1 is in flight while
// chunk c is processed.  loaded() runs once every TMEM read of the row has completed (before the last chunk's body).
template <int NCHUNK, class Body, class Loaded>
__device__ __forceinline__ void chunk_loop(uint32_t taddr, Body&& body, Loaded&& loaded) {
  uint32_t va[32], vb[32];
  tmem_ld32(taddr, va);
  tmem_ld_wait();
  if (NCHUNK == 1) loaded();
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    uint32_t (&cur)[32] = (c & 1) ? vb : va;
    uint32_t (&nxt)[32] = (c & 1) ? va : vb;
    if (c + 1 < NCHUNK) tmem_ld32(taddr + (c + 1) * 32, nxt);
    body(c, cur);
    if (c + 1 < NCHUNK) tmem_ld_wait();
    if (c + 2 == NCHUNK) loaded();
  }
}

// The body of one CTA of a pair: `pid` of `npid` pairs in this role (the plain kernels have one role; the merged
// dx+dW kernel gives the first pairs the DX role and the rest the DW role).
template <int MODE, int V>
__device__ __forceinline__ void tc_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& a, const int64_t pid,
                                        const int64_t npid, uint8_t* smem_raw) {
  constexpr int STAGES = mode_stages(MODE);
  constexpr int STAGE_BYTES = mode_stage_bytes(MODE);
  constexpr bool AS = mode_astat(MODE);              // x^ tile resident in smem, only w^ streams
  constexpr int A_IN_STAGE = AS ? 0 : A_STAGE_BYTES;
  constexpr int GW = BN / 2;                         // B columns (of one 256-wide UMMA group) held by this CTA
  constexpr int BNT = mode_bn(MODE);                 // accumulator columns of one tile (256; DX, DW: 512)
  constexpr int NBUF = mode_nbuf(MODE);
  constexpr bool A_MN = (MODE == MODE_DW);
  constexpr bool B_MN = (MODE == MODE_DX || MODE == MODE_DW);
  const int rank = (int)cluster_ctarank();
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;           // SWIZZLE_128B needs 1024 B alignment
  const uint32_t ares_base = smem_base;                             // AS: resident A, 8 k-blocks x 16 KB
  const uint32_t tiles_base = smem_base + (AS ? A_RESIDENT_BYTES : 0);
  uint8_t* tiles_ptr = smem_raw + (tiles_base - raw_addr);
  constexpr int W_BYTES = (MODE == MODE_DW) ? W_TILE_BYTES : 0;
  const uint32_t wtile_base = tiles_base + STAGES * STAGE_BYTES;    // DW: w^ tile, 4 boxes [128 classes][64 d] (1024 B aligned)
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles_ptr + STAGES * STAGE_BYTES + W_BYTES);
  uint8_t* stg_all = tiles_ptr + STAGES * STAGE_BYTES + W_BYTES + 256;   // output staging (FWDS / BWD_G / DW)
  const uint32_t bar_full = smem_u32(bars);                         // [STAGES]
  const uint32_t bar_empty = bar_full + 8 * MAX_STAGES;             // [STAGES]
  const uint32_t bar_tfull = bar_empty + 8 * MAX_STAGES;            // [2]
  const uint32_t bar_tempty = bar_tfull + 16;                       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);
  const uint32_t bar_afull = bar_full + 8 * (2 * MAX_STAGES + 5);   // AS: resident A landed
  const uint32_t bar_afree = bar_afull + 8;                         // AS: every MMA that read the resident A retired
  const uint32_t bar_rdone = bar_afree + 8;                         // DX side pass: [STAGES] the epilogue warps read the A tile
  const uint32_t bar_wfull = bar_rdone + 8 * MAX_STAGES;            // DW: w^ tile landed
  const uint32_t bar_wempty = bar_wfull + 8;                        // DW: the epilogue warps are done with the w^ tile

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (MODE == MODE_DW) prefetch_tmap(&a.tmW);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 2);                   // leader's expect_tx arrive + the peer producer's arrive
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, NUM_EPI_WARPS * 2);   // one arrive per epilogue warp of both CTAs
    }
    mbar_init(bar_afull, 2);
    mbar_init(bar_afree, 1);
    for (int s = 0; s < STAGES; ++s) mbar_init(bar_rdone + 8 * s, NUM_EPI_WARPS);
    mbar_init(bar_wfull, 1);
    mbar_init(bar_wempty, NUM_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  // barriers, TMEM and tensor-map prefetch above touch nothing a predecessor writes: with programmatic dependent launch
  // they overlap the previous kernel's tail; every global / TMA access of every role comes after this point
  mh_pdl_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer (both CTAs of the pair) ===============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      TileLoop<MODE> tl;
      tl.init(a, pid, npid);
      Work w;
      int res_m = -1;
      uint32_t tile_j = 0;
      // L2 priorities.  DX: the stash / G tile is read by exactly one pair (evict first; in the merged kernel the dW
      // role still wants it: normal), w^ by every row tile of the split (evict last).  DW: x^ is the 1 MB every tile
      // re-reads (evict last).
      const uint64_t pol_a = (MODE == MODE_DX && !a.prog) ? l2_policy_evict_first() : 0;
      const uint64_t pol_b = (MODE == MODE_DX || MODE == MODE_DW) ? l2_policy_evict_last() : 0;
      while (tl.next(a, rank, w)) {
        if (AS && w.m_tile != res_m) {
          // (re)load the resident x^ tile: wait until every MMA of the previous tiles has retired
          if (tile_j > 0) mbar_wait(bar_afree, (tile_j - 1) & 1);
          if (rank == 0) mbar_expect_tx(bar_afull, 2 * A_RESIDENT_BYTES); else mbar_arrive_cluster(bar_afull, 0);
          for (int kb = 0; kb < MH_D / BK; ++kb)
            tma_load_2d_2sm(ares_base + kb * A_STAGE_BYTES, &tmA, bar_afull, kb * BK, w.m0);
          res_m = w.m_tile;
        }
        ++tile_j;
        if (AS && a.w_ready) {
          // merged prologue+forward kernel: this class tile of w^ is being written by the prologue role of the same
          // launch; wait for all 256 rows (release / acquire on the tile's counter), then order the generic-proxy writes
          // before this thread's TMA (async-proxy) reads
          if (pid == 0 && rank == 0) st_relaxed_gpu(a.fwd_front, w.n_tile);
          const int* rdy = a.w_ready + w.n_tile;
          if (ld_acquire_gpu(rdy) < BN) {
            const long long t0 = clock64();
            while (ld_acquire_gpu(rdy) < BN) {
              if (clock64() - t0 > 4000000000LL) __trap();
            }
          }
          asm volatile("fence.proxy.async.global;" ::: "memory");
        }
        if (MODE == MODE_DW && a.prog) {
          // merged dx+dW kernel: stay within PROG_AHEAD class tiles of the dx front so that the stash and w^ tiles the dx
          // pairs fetched are still in L2 (the role that is behind never waits: no deadlock; the leader publishes)
          const int ct = w.m0 / BMT;
          if (pid == 0 && rank == 0) st_relaxed_gpu(a.prog + 1, ct);
          if (ct > ld_acquire_gpu(a.prog) + a.prog_ahead) {
            const long long t0 = clock64();
            while (ct > ld_acquire_gpu(a.prog) + a.prog_ahead) {
              __nanosleep(256);
              if (clock64() - t0 > 4000000000LL) __trap();
            }
          }
        }
        for (int kb = w.kb0; kb < w.kb1; kb = kb_next(w, kb)) {
          if (MODE == MODE_DX && a.prog && w.kb_chunk && (kb - w.kb0) % w.kb_chunk == 0) {
            const int ct = kb / w.kb_chunk;                    // 256-class tile index (dx_chunk = 4 k-blocks of 64 classes)
            if (pid == 0 && rank == 0) st_relaxed_gpu(a.prog, ct);
            if (ct > ld_acquire_gpu(a.prog + 1) + a.prog_ahead) {
              const long long t0 = clock64();
              while (ct > ld_acquire_gpu(a.prog + 1) + a.prog_ahead) {
                __nanosleep(256);
                if (clock64() - t0 > 4000000000LL) __trap();
              }
            }
          }
          if (MODE == MODE_DX && a.dx_sync && !w.kb_chunk && ((kb - w.kb0) & (DX_SYNC_KB - 1)) == 0) {
            // lockstep among the 2 * m_tiles CTAs that stream the same w^ k-blocks (one split, all row tiles): nobody
            // starts k-block group g before everybody has issued group g-1, so a w^ tile fetched from HBM for one row
            // tile is still in L2 for the others (ncu: 7.6 GB read for 6.15 GB algorithmic without it).  Single wave only
            // (host-checked: every tile is resident), bounded spin.
            int* ctr = a.dx_sync + w.split;
            const int target = ((kb - w.kb0) / DX_SYNC_KB) * 2 * a.m_tiles;
            red_release_gpu_add(ctr, 1);
            if (ld_acquire_gpu(ctr) < target) {
              const long long t0 = clock64();
              while (ld_acquire_gpu(ctr) < target) {
                if (clock64() - t0 > 4000000000LL) __trap();
              }
            }
          }
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          if (MODE == MODE_DX && a.rho) mbar_wait(bar_rdone + 8 * stage, phase ^ 1);   // side pass has read the A tile
          const uint32_t sa = tiles_base + stage * STAGE_BYTES;
          const uint32_t sb = sa + A_IN_STAGE;
          const uint32_t fb = bar_full + 8 * stage;
          // the bytes of both CTAs land on the leader's barrier
          if (rank == 0) mbar_expect_tx(fb, 2 * STAGE_BYTES); else mbar_arrive_cluster(fb, 0);
          if (MODE == MODE_DX) {
            // A = G, class-tiled [C_pad/128][B_pad][128]: k-block kb = classes 64kb.. -> slab kb/2, columns (kb&1)*64
            tma_load_2d_2sm_hint(sa, &tmA, fb, (kb & 1) * 64, (kb >> 1) * (int)a.B_pad + w.m0, pol_a);   // box [64 k][128 rows]
          } else if (MODE == MODE_DW) {
            // A = G^T from the class-tiled G: this CTA's 128 classes are slab m0/128; boxes [64 classes][64 rows]
#pragma unroll
            for (int bx = 0; bx < BM / 64; ++bx)
              tma_load_2d_2sm(sa + bx * 8192, &tmA, fb, 64 * bx, (w.m0 / 128) * (int)a.B_pad + kb * BK);
          }
          if (!B_MN) {
            tma_load_2d_2sm(sb, &tmB, fb, kb * BK, w.n0 + rank * GW);                        // box [64 k][128 rows]
          } else {
#pragma unroll
            for (int nh = 0; nh < BNT / BN; ++nh)
#pragma unroll
              for (int bx = 0; bx < GW / 64; ++bx)
                tma_load_2d_2sm_hint(sb + (nh * (GW / 64) + bx) * 8192, &tmB, fb, w.n0 + nh * BN + rank * GW + 64 * bx, kb * BK,
                                     pol_b);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (MODE == MODE_DW && !a.raw_dw) {
          // w^ tile for this tile's epilogue.  Issued after the k-loop loads so that waiting for the previous tile's
          // epilogue (which overlaps this tile's MMAs) never holds back the operand pipeline.
          if (tile_j > 1) mbar_wait(bar_wempty, (tile_j - 2) & 1);
          mbar_expect_tx(bar_wfull, W_TILE_BYTES);
#pragma unroll
          for (int bx = 0; bx < BN / 64; ++bx)
            tma_load_2d_local(wtile_base + bx * (BM * 128), &a.tmW, bar_wfull, w.n0 + 64 * bx, w.m0);
        }
      }
      // this role has issued all its loads: release the other role from the cross-role throttle
      if ((MODE == MODE_DX || MODE == MODE_DW) && a.prog && pid == 0 && rank == 0)
        st_relaxed_gpu(a.prog + (MODE == MODE_DW ? 1 : 0), 0x3fffffff);
      if (AS && a.w_ready && pid == 0 && rank == 0) st_relaxed_gpu(a.fwd_front, 0x3fffffff);
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA only) ===============================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc(A_MN ? 1 : 0, B_MN ? 1 : 0, BMT, BN);
      uint32_t stage = 0, phase = 0;
      uint32_t it = 0;
      TileLoop<MODE> tl;
      tl.init(a, pid, npid);
      Work w;
      int res_m = -1;
      uint32_t aphase = 0;
      for (; tl.next(a, rank, w); ++it) {
        const uint32_t buf = it % NBUF, bphase = (it / NBUF) & 1;
        mbar_wait(bar_tempty + 8 * buf, bphase ^ 1);                       // epilogues drained this accumulator
        if (AS && w.m_tile != res_m) {
          mbar_wait(bar_afull, aphase);
          aphase ^= 1;
          res_m = w.m_tile;
        }
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BN;
        for (int kb = w.kb0; kb < w.kb1; kb = kb_next(w, kb)) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = AS ? (ares_base + kb * A_STAGE_BYTES) : (tiles_base + stage * STAGE_BYTES);
          const uint32_t sb = tiles_base + stage * STAGE_BYTES + A_IN_STAGE;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = A_MN ? desc_mnmajor(sa, k) : desc_kmajor(sa, k);
            const uint32_t acc = (kb > w.kb0 || k > 0) ? 1u : 0u;
#pragma unroll
            for (int nh = 0; nh < BNT / BN; ++nh) {                       // DX, DW: two N=256 groups of the 512-wide tile
              const uint32_t sbh = sb + nh * (GW / 64) * 8192;
              const uint64_t db = B_MN ? desc_mnmajor(sbh, k) : desc_kmajor(sbh, k);
              umma_bf16_2sm(tmem_d + nh * BN, da, db, idesc, acc);
            }
          }
          umma_commit_2sm(bar_empty + 8 * stage);      // frees the smem stage in both CTAs when the MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(bar_tfull + 8 * buf);          // accumulator ready for the epilogues
        if (AS) umma_commit_2sm(bar_afree);
      }
    }
  } else if (warp >= EPI_WARP0) {
    // =============================== epilogue ===============================
    // 8 warps: warp%4 selects the TMEM lane quarter (hardware rule), (warp-4)/4 the column half.
    const int q = warp & 3;
    const int half = (warp - EPI_WARP0) >> 2;
    const int r = q * 32 + lane;                    // row of the tile owned by this thread
    constexpr int NCHUNK = (BNT / 2) / 32;          // 32-column chunks per warp (4; DX, DW: 8)
    const MhParams& p = a.p;
    const float ha = (V == V_CURR) ? a.state[4] : p.hard_a;
    const float hb = p.hard_b, lo = p.lo, hi = p.hi;
    uint8_t* stg = stg_all + (warp - EPI_WARP0) * STG_WARP_BYTES;      // this warp's staging tile
    uint32_t it = 0;
    uint32_t side_stage = 0, side_phase = 0;          // DX side pass: walks the smem stages like the producer
    // A-stationary modes: the thread keeps its row for as long as the pair keeps its row tile, so the row terms are
    // loaded once and the softmax statistics accumulate in registers across class tiles (one record per pair, row and
    // column half instead of one per tile).
    int res_m = -1, res_y = -1;
    int64_t res_row = 0;
    RowCtx rc{};
    FwdAcc acc{-INFINITY, 0.f, 0.f, 0.f, 0};
    auto flush_stats = [&]() {
      if (MODE == MODE_FWD || MODE == MODE_FWDS) {
        float* sp = a.stats_tiles + ((int64_t)pid * 2 + half) * MH_ST_PLANES * a.B_pad;
        sp[MH_ST_M * a.B_pad + res_row] = acc.m;
        sp[MH_ST_L * a.B_pad + res_row] = acc.l;
        sp[MH_ST_CNT * a.B_pad + res_row] = (float)acc.cnt + acc.cntf;
        sp[MH_ST_EZ * a.B_pad + res_row] = (V == V_SPHERE) ? acc.ez : 0.f;
      }
    };
    TileLoop<MODE> tl;
    tl.init(a, pid, npid);
    Work w;
    for (; tl.next(a, rank, w); ++it) {
      const uint32_t buf = it % NBUF, bphase = (it / NBUF) & 1;
      const int64_t row = (int64_t)w.m0 + r;
      const int cbase = half * (BNT / 2);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + cbase;
      auto release = [&]() {
        // all TMEM reads of this accumulator are complete -> hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(bar_tempty + 8 * buf, 0);
      };
      auto no_op = [&]() {};

      if (AS) {
        const bool fix = (MODE == MODE_FWDS) || (MODE == MODE_FWD && V != V_SPHERE && a.fixref);
        if (w.m_tile != res_m) {
          // a new resident row tile: flush the statistics of the previous one, (re)load this thread's row terms
          if (res_m >= 0) flush_stats();
          res_m = w.m_tile; res_row = row;
          rc.valid = row < a.B;
          rc.scale = a.rowp[MH_RP_SCALE * a.ldp + row];
          rc.scale2 = rc.scale * MH_LOG2E;
          rc.thr = a.rowp[MH_RP_THR * a.ldp + row];
          rc.t = a.rowp[MH_RP_T * a.ldp + row];
          rc.zt2 = a.rowp[MH_RP_ZT * a.ldp + row] * MH_LOG2E;
          rc.dzt = a.rowp[MH_RP_DZT * a.ldp + row];
          rc.nref2 = 102.f - rc.scale2 * a.umax;
          rc.ntbig = -rc.t * CNT_BIG;
          rc.lse2 = (MODE == MODE_BWD_G && rc.valid) ? a.lse2[row] : 0.f;
          res_y = a.label_local[row];
          acc = FwdAcc{fix ? -rc.nref2 : -INFINITY, 0.f, 0.f, 0.f, 0};
        }
        rc.tcol = (res_y >= w.n0 && res_y < w.n0 + BN) ? res_y - w.n0 : -1;
        const int nvalid = (int)min((int64_t)BN, a.C - (int64_t)w.n0);   // valid class columns in this tile
        mbar_wait(bar_tfull + 8 * buf, bphase);
        tc_fence_after();
        chunk_loop<NCHUNK>(taddr, [&](int c, uint32_t (&cur)[32]) {
          const int col0 = cbase + c * 32;
          uint32_t pk[16];
          if (MODE == MODE_FWD) {
            if (V != V_SPHERE && fix) fwd_chunk_fix<V == V_SPHERE ? V_CLAMP : V, false>(cur, col0, nvalid, rc, lo, hi, ha, hb, acc, pk);
            else fwd_chunk<V>(cur, col0, nvalid, rc, lo, hi, ha, hb, acc);
          } else {
            if (MODE == MODE_FWDS) {
              fwd_chunk_fix<V, true>(cur, col0, nvalid, rc, lo, hi, ha, hb, acc, pk);
            } else {
              float qv[32];
              bwd_chunk<V>(cur, col0, nvalid, rc, lo, hi, ha, hb, pk, qv);
              if (a.rsum) {
                const float cs = warp_colsum32(qv, lane);           // lane L: sum over this warp's 32 rows, column col0+L
                atomicAdd(a.rsum + w.n0 + col0 + lane, cs);
              }
            }
            stage_half_row(stg, lane, c & 1, pk);
            if (c & 1) {
              // two chunks staged = [32 rows][64 classes = 128 B] -> global as full lines.  The B x C buffer is
              // class-tiled [C_pad/128][B_pad][128]: this warp's 128 columns are one slab.
              __syncwarp();
              __nv_bfloat16* obase = a.G + (((int64_t)(w.n0 + cbase) / 128) * a.B_pad + w.m0 + q * 32) * 128 + (c - 1) * 32;
              flush_tile(stg, lane, reinterpret_cast<uint8_t*>(obase), 256, 32);
              __syncwarp();
            }
          }
        }, release);
      } else if (MODE == MODE_DX) {
        if (a.rho) {
          // ---- side pass (stash mode): r_j += sum_i rho_i E'_ij cos_ij over this CTA's 128 rows, k-block by k-block.
          // An A tile is [128 rows][64 classes] (128 B rows, 16 B chunks XOR-swizzled with row & 7).  Warp ew owns the
          // 8 classes of chunk ew, lane L the rows L, L+32, L+64, L+96: 4 conflict-free 16 B loads per k-block, taken
          // once the tile's MMAs have retired (bar_empty) and released to the producer at once (bar_rdone); the row
          // sums are then reduced with shuffles -- no shared-memory traffic, no CTA barrier.  cos_ij is recovered from
          // the stash: E' = exp2(s2 cos - ref2)  =>  cos = log2(E') / s2 + ref2 / s2  (E' = 0: target / clamped / pad).
          const int ew = warp - EPI_WARP0;
          float rh[4], ucut[4];
          // MV-Softmax: a stashed value is e * 1 with u = c (c <= thr_i) or e * a with u = a*c + b (c > thr_i).  Read as
          // u_easy = log2(E')/s2 + kappa the two cases fall into disjoint ranges, (-inf, thr_i] and
          // (a*thr_i + b + log2(a)/s2, +inf): ucut is the midpoint of the gap, so the branch is recovered exactly.
          const float lg_ha = a.side_mv ? log2f(a.side_ha) : 0.f;
          const float kappa_h = a.side_kappa - lg_ha * a.side_inv_s2, ha_inv = a.side_mv ? 1.f / a.side_ha : 1.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            rh[j] = a.rho[w.m0 + lane + 32 * j];
            const float thr = a.side_mv ? a.rowp[MH_RP_THR * a.ldp + w.m0 + lane + 32 * j] : 0.f;
            ucut[j] = thr + 0.5f * ((a.side_ha - 1.f) * thr + a.side_hb + lg_ha * a.side_inv_s2);
          }
          for (int kb = w.kb0; kb < w.kb1; kb = kb_next(w, kb)) {
            mbar_wait(bar_empty + 8 * side_stage, side_phase);
            const uint32_t sa = tiles_base + side_stage * STAGE_BYTES;
            uint4 q4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int ar = lane + 32 * j;
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(q4[j].x), "=r"(q4[j].y), "=r"(q4[j].z), "=r"(q4[j].w)
                           : "r"(sa + ar * 128 + ((ew ^ (ar & 7)) << 4)));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_local(bar_rdone + 8 * side_stage);
            float acc8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc8[e] = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t ww[4] = {q4[j].x, q4[j].y, q4[j].z, q4[j].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float v0 = __uint_as_float(ww[e] << 16), v1 = __uint_as_float(ww[e] & 0xffff0000u);
                const float l0 = fmaxf(lg2(v0), -200.f), l1 = fmaxf(lg2(v1), -200.f);
                float c0 = fmaf(l0, a.side_inv_s2, a.side_kappa);
                float c1 = fmaf(l1, a.side_inv_s2, a.side_kappa);
                if (a.side_mv) {
                  if (c0 > ucut[j]) c0 = (fmaf(l0, a.side_inv_s2, kappa_h) - a.side_hb) * ha_inv;
                  if (c1 > ucut[j]) c1 = (fmaf(l1, a.side_inv_s2, kappa_h) - a.side_hb) * ha_inv;
                }
                acc8[2 * e] = fmaf(rh[j] * v0, c0, acc8[2 * e]);
                acc8[2 * e + 1] = fmaf(rh[j] * v1, c1, acc8[2 * e + 1]);
              }
            }
            // transposed butterfly over the 32 lanes: 8 -> 4 -> 2 -> 1 values per lane, then two plain steps
#pragma unroll
            for (int sft = 4; sft >= 1; sft >>= 1) {
              const bool upper = (lane & sft) != 0;
#pragma unroll
              for (int e = 0; e < sft; ++e) {
                const float send = upper ? acc8[e] : acc8[e + sft];
                const float keep = upper ? acc8[e + sft] : acc8[e];
                acc8[e] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
              }
            }
            float t = acc8[0];
            t += __shfl_xor_sync(0xffffffffu, t, 8);
            t += __shfl_xor_sync(0xffffffffu, t, 16);
            // lane L (< 8) now holds class bit-reversal-free index: value e = (L&4 ? 4:0)|(L&2 ? 2:0)|(L&1 ? 1:0)
            // every (128-row block, class) pair is produced exactly once: plain store, no atomics, no memset
            if (lane < 8) a.rsum[(int64_t)(w.m0 / BM) * a.C_pad + (int64_t)kb * BK + ew * 8 + lane] = t;
            if (++side_stage == STAGES) { side_stage = 0; side_phase ^= 1; }
          }
        }
        mbar_wait_long(bar_tfull + 8 * buf, bphase);
        tc_fence_after();
        chunk_loop<NCHUNK>(taddr, [&](int c, uint32_t (&cur)[32]) {
          float* dst = a.out + (int64_t)w.split * a.out_split_stride + row * MH_D + cbase + c * 32;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            MH_STCS(reinterpret_cast<uint4*>(dst) + k, make_uint4(cur[4 * k], cur[4 * k + 1], cur[4 * k + 2], cur[4 * k + 3]));
        }, release);
      } else {
        // ---- DW: this thread owns class `row` and 128 of the tile's 256 d columns ----
        const bool raw = a.raw_dw != 0;
        const bool ok = raw || row < a.C;
        const bool selfp = !raw && a.rpart != nullptr;
        float rj = 0.f, coef = 1.f;
        if (!raw && row < a.C) {
          if (!selfp)
            for (int pp = 0; pp < a.rsum_parts; ++pp) rj += a.rsum[(int64_t)pp * a.C_pad + row];   // fixed order
          coef = a.gscal[0] * a.inv_norm[row];
        }
        if (!raw && row >= a.C) coef = 0.f;
        const int rows_ok = raw ? 32 : (int)max((int64_t)0, min((int64_t)32, a.C - ((int64_t)w.m0 + q * 32)));
        const int64_t opitch = raw ? (int64_t)MH_D : a.ld;
        mbar_wait(bar_tfull + 8 * buf, bphase);
        tc_fence_after();
        if (!raw) mbar_wait(bar_wfull, it & 1);            // this tile's w^ [128 classes][256 d] is in shared memory
        if (selfp) {
          // ---- self-projection: r_j = w^_j . dw^_j taken from the accumulators themselves.  This thread owns class `row`
          // and 128 of the 512 d columns (d half of the tile x column half of the warp); the four partial dots of a class
          // live in two CTA pairs (the tiles 2k / 2k+1 = the two d halves of one class tile run on neighbouring pairs at
          // the same time).  Every partial is posted to global memory BEFORE anything waits, the wait is bounded, and
          // the four partials are summed in a fixed order: no atomics on data, bit-reproducible, no deadlock by
          // construction (a tile's post depends only on its own pair reaching it).
          float p = 0.f;
          chunk_loop<NCHUNK>(taddr, [&](int c, uint32_t (&cur)[32]) {
            const int dcol = cbase + c * 32;
            const uint32_t wrow_s = wtile_base + (dcol >> 6) * (BM * 128) + r * 128;
            const int ci0 = (dcol & 63) >> 3;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              uint4 wq;
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(wq.x), "=r"(wq.y), "=r"(wq.z), "=r"(wq.w)
                           : "r"(wrow_s + (((ci0 + k4) ^ (r & 7)) << 4)));
              const uint32_t ww[4] = {wq.x, wq.y, wq.z, wq.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 wf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ww[e]));
                const int k = k4 * 8 + e * 2;
                p = fmaf(__uint_as_float(cur[k]), wf.x, p);
                p = fmaf(__uint_as_float(cur[k + 1]), wf.y, p);
              }
            }
          }, no_op);
          const int plane = (w.n0 / BN) * 2 + half;
          __stcg(a.rpart + (int64_t)plane * a.C_pad + row, p);
          __threadfence();
          epi_bar_sync();                                       // all 256 partials of this CTA are posted and fenced
          int* flag = a.rflag + (w.m0 / BM);
          if (warp == EPI_WARP0 && lane == 0) {
            red_release_gpu_add(flag, 1);
            if (ld_acquire_gpu(flag) < 2) {                     // the CTA holding the other d half of these 128 classes
              const long long t0 = clock64();
              while (ld_acquire_gpu(flag) < 2) {
                if (clock64() - t0 > 4000000000LL) __trap();
              }
            }
          }
          epi_bar_sync();
          rj = __ldcg(a.rpart + row) + __ldcg(a.rpart + a.C_pad + row) + __ldcg(a.rpart + 2 * a.C_pad + row) +
               __ldcg(a.rpart + 3 * a.C_pad + row);
        }
        chunk_loop<NCHUNK>(taddr, [&](int c, uint32_t (&cur)[32]) {
          float o[32];
          if (raw) {
#pragma unroll
            for (int k = 0; k < 32; ++k) o[k] = __uint_as_float(cur[k]);
          } else {
            // w^ row r, d columns cbase + 32c ..: box (cbase + 32c) / 64, 16 B chunks XOR-swizzled with r & 7
            const int dcol = cbase + c * 32;
            const uint32_t wrow_s = wtile_base + (dcol >> 6) * (BM * 128) + r * 128;
            const int ci0 = (dcol & 63) >> 3;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              uint4 wq;
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(wq.x), "=r"(wq.y), "=r"(wq.z), "=r"(wq.w)
                           : "r"(wrow_s + (((ci0 + k4) ^ (r & 7)) << 4)));
              const uint32_t ww[4] = {wq.x, wq.y, wq.z, wq.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 wf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ww[e]));
                const int k = k4 * 8 + e * 2;
                o[k] = (__uint_as_float(cur[k]) - wf.x * rj) * coef;
                o[k + 1] = (__uint_as_float(cur[k + 1]) - wf.y * rj) * coef;
              }
            }
            if (c == NCHUNK - 1) {
              __syncwarp();
              if (lane == 0) mbar_arrive_local(bar_wempty);      // producer may fetch the next tile's w^
            }
          }
          if (raw || a.layout == MH_LAYOUT_CD) {
            // stage [32 rows][32 fp32 = 128 B], then full-line stores: 8 lanes per row, 4 rows per instruction
#pragma unroll
            for (int k = 0; k < 8; ++k)
              *reinterpret_cast<float4*>(stg + lane * 128 + ((k ^ (lane & 7)) * 16)) =
                  make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
            __syncwarp();
            float* obase = a.out + ((int64_t)w.m0 + q * 32) * opitch + w.n0 + cbase + c * 32;
            flush_tile(stg, lane, reinterpret_cast<uint8_t*>(obase), opitch * 4, rows_ok);
            __syncwarp();
          } else if (ok) {
            // parameter layout [D, C]: for a fixed d the 32 lanes are 32 consecutive classes -> coalesced directly
#pragma unroll
            for (int k = 0; k < 32; ++k) MH_STCS(a.out + (int64_t)(w.n0 + cbase + c * 32 + k) * a.ld + row, o[k]);
          }
        }, release);
      }
    }
    if (AS && res_m >= 0) flush_stats();
  }
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();                                     // the peer's smem / TMEM stay valid until both are done
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

template <int MODE, int V>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
          const __grid_constant__ TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  if (a.gate) {                       // gated launch (guarded stash): the flag is a predecessor's output, decide first
    mh_pdl_sync();
    if (gate_closed(a.gate, a.gate_on)) return;
  }
  tc_body<MODE, V>(tmA, tmB, a, blockIdx.x >> 1, gridDim.x >> 1, smem_raw);   // tile-scheduling unit: CTA pair
}

// Merged backward: pairs [0, n_dx) compute dx^ partials (DX role, interleaved class chunks), the remaining pairs compute
// dW (DW role, self-projecting).  Both roles walk the classes in the same direction and keep within PROG_AHEAD class
// tiles of each other (TcArgs::prog), so the stash and w^ tiles are fetched from HBM once and hit in L2 for the other
// role: 6.15 GB less DRAM traffic per step than two back-to-back kernels.  All CTAs are co-resident (grid <= #SMs, one
// CTA per SM), which the cross-role throttle and the dW partner exchange rely on.
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_kernel_dxdw(const __grid_constant__ CUtensorMap tmA_dx, const __grid_constant__ CUtensorMap tmB_dx,
               const __grid_constant__ TcArgs a_dx, const __grid_constant__ CUtensorMap tmA_dw,
               const __grid_constant__ CUtensorMap tmB_dw, const __grid_constant__ TcArgs a_dw, const int n_dx) {
  extern __shared__ uint8_t smem_raw[];
  const int64_t pair = blockIdx.x >> 1, pairs = gridDim.x >> 1;
  if (pair < n_dx) tc_body<MODE_DX, V_NONE>(tmA_dx, tmB_dx, a_dx, pair, n_dx, smem_raw);
  else tc_body<MODE_DW, V_NONE>(tmA_dw, tmB_dw, a_dw, pair - n_dx, pairs - n_dx, smem_raw);
}

// ---- merged W prologue + forward -------------------------------------------------------------------------
// The W prologue ([C, 512] layout) as a ROLE of the forward launch: a few CTA pairs normalise the class centres tile by
// tile (fp32 W -> bf16 w^ + 1/|w|), the other pairs run the fused forward and pick every w^ tile up from L2 right after
// it was written.  The HBM-bound prologue (0.87 ms alone at cfg4) disappears under the tensor-bound forward and the
// forward's 2.05 GB read of w^ never reaches DRAM.  Protocol: per 256-class tile a counter of finished rows
// (red.release after the rows' stores; the forward producer acquires it, then fence.proxy.async before its TMA loads);
// the prologue role stays at most PW_AHEAD tiles ahead of the forward front (L2 residency), and only the role that is
// ahead ever waits.  Bounded spins, all CTAs resident (grid = #SMs).
constexpr int PW_AHEAD = 96;
struct PwArgs {
  const float* W;
  int64_t C, C_pad, ld;
  __nv_bfloat16* what;
  float* inv_norm;
  int* ready;                  // [C_pad / 256]
  const int* fwd_front;
};

// Prologue role.  One warp normalises PW_G = 4 consecutive classes per step; the fp32 rows arrive by TMA into a per-warp
// double-buffered shared-memory slot (2 x 8 KB per warp, 12 warps: 192 KB), so the loads of the next step are in flight
// while this step computes, stores and publishes - the release fence that publishes (it has to wait for the stores) never
// stalls the load stream.  Same arithmetic and summation order as
// prologue_w_cd_kernel (prologue.cu): w^ and inv_norm are bit-identical to the stand-alone prologue.
constexpr int PW_G = 4;                              // classes per warp step
constexpr int PW_SLOT = PW_G * MH_D * 4;             // 8 KB of fp32 rows
constexpr int PW_WARPS = NUM_THREADS / 32;
constexpr int PW_BATCH = 4;                          // warp steps per publish
static_assert(2 * PW_WARPS * PW_SLOT + 1024 <= 200 * 1024, "prologue-role staging must fit in shared memory");

__device__ __forceinline__ void pw_role(const PwArgs& p, const CUtensorMap* tmW32, int cta, int ncta, uint8_t* smem_raw) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t slots = base + 256;                                     // [PW_WARPS][2][PW_SLOT]
  const uint32_t bar0 = base + warp * 16;                                // two mbarriers per warp
  uint8_t* slots_ptr = smem_raw + (slots - smem_u32(smem_raw));
  if (lane == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_barrier_init();
  }
  __syncwarp();
  const int64_t groups = p.C_pad / PW_G;
  const int64_t stride = (int64_t)ncta * PW_WARPS;
  // the fp32 rows arrive by 2-D TMA (the tensor engine's tiled path; 1-D bulk copies sustained only ~30 GB/s per SM):
  // two boxes of [PW_G rows][256 columns] per group, rows beyond C read as zeros (out-of-bounds fill)
  auto issue = [&](int64_t g, int slot) {            // lane 0: start the copies of group g into `slot`
    const int row0 = (int)(g * PW_G);
    const uint32_t bar = bar0 + 8 * slot;
    const uint32_t dst = slots + (uint32_t)((warp * 2 + slot) * PW_SLOT);
    mbar_expect_tx(bar, PW_SLOT);
    tma_load_2d_local(dst, tmW32, bar, 0, row0);
    tma_load_2d_local(dst + PW_SLOT / 2, tmW32, bar, 256, row0);
  };
  int64_t g = (int64_t)cta * PW_WARPS + warp;
  int pend[PW_BATCH], npend = 0;
  uint32_t ph[2] = {0u, 0u};
  int slot = 0;
  // stay at most PW_AHEAD class tiles ahead of the forward front.  The front is re-read only when the cached value says
  // "too far ahead" or every 8th step (a global load per step in lane 0's issue path cost more than the copies)
  int front = 0, since = 8;
  auto throttle = [&](int64_t gg) {
    const int tile = (int)(gg * PW_G / BN);
    if (++since >= 8) { front = ld_acquire_gpu(p.fwd_front); since = 0; }
    if (tile > front + PW_AHEAD) {
      const long long t0 = clock64();
      while (tile > (front = ld_acquire_gpu(p.fwd_front)) + PW_AHEAD) {
        if (clock64() - t0 > 4000000000LL) __trap();
      }
      since = 0;
    }
  };
  if (g < groups && lane == 0) { throttle(g); issue(g, 0); }
  for (; g < groups; g += stride) {
    const int64_t gn = g + stride;
    if (gn < groups && lane == 0) { throttle(gn); issue(gn, slot ^ 1); }
    mbar_wait(bar0 + 8 * slot, ph[slot]);
    ph[slot] ^= 1u;
    const float4* src = reinterpret_cast<const float4*>(slots_ptr + (warp * 2 + slot) * PW_SLOT);
    const int64_t row0 = g * PW_G;
    // the PW_G classes of a step are independent: keep their dependency chains (square sums, warp reductions, the
    // reciprocal) interleaved - with 12 warps per SM the role is latency-bound, not bandwidth-bound, otherwise
    float4 v[PW_G][4];
    float ss[PW_G], inv[PW_G];
#pragma unroll
    for (int r = 0; r < PW_G; ++r) {
      ss[r] = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) v[r][k] = src[(k >> 1) * (PW_SLOT / 32) + r * 64 + lane + 32 * (k & 1)];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int r = 0; r < PW_G; ++r)
        ss[r] += v[r][k].x * v[r][k].x + v[r][k].y * v[r][k].y + v[r][k].z * v[r][k].z + v[r][k].w * v[r][k].w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < PW_G; ++r) ss[r] += __shfl_xor_sync(0xffffffffu, ss[r], o);
#pragma unroll
    for (int r = 0; r < PW_G; ++r) inv[r] = 1.f / fmaxf(sqrtf(ss[r]), 1e-12f);
#pragma unroll
    for (int r = 0; r < PW_G; ++r) {
      const int64_t row = row0 + r;
      uint2* dst = reinterpret_cast<uint2*>(p.what + row * MH_D);
      if (lane == 0 && row < p.C) p.inv_norm[row] = inv[r];
#pragma unroll
      for (int k = 0; k < 4; ++k) {                  // rows >= C arrive as zeros (TMA out-of-bounds fill) and stay zero
        const float4 o = make_float4(v[r][k].x * inv[r], v[r][k].y * inv[r], v[r][k].z * inv[r], v[r][k].w * inv[r]);
        __nv_bfloat162 p0 = __floats2bfloat162_rn(o.x, o.y), p1 = __floats2bfloat162_rn(o.z, o.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&p0);
        pk.y = *reinterpret_cast<uint32_t*>(&p1);
        dst[lane + 32 * k] = pk;
      }
    }
    // publish every PW_BATCH steps: one release fence (it waits for the steps' stores to be performed) covers them all;
    // every lane's stores are ordered before lane 0's fence by the warp barrier (cumulativity)
    pend[npend++] = (int)(row0 / BN);
    if (npend == PW_BATCH) {
      __syncwarp();
      if (lane == 0) {
        __threadfence();
#pragma unroll
        for (int i = 0; i < PW_BATCH; ++i) red_relaxed_gpu_add(p.ready + pend[i], PW_G);
      }
      npend = 0;
    }
    slot ^= 1;
  }
  __syncwarp();
  if (lane == 0 && npend > 0) {
    __threadfence();
    for (int i = 0; i < npend; ++i) red_relaxed_gpu_add(p.ready + pend[i], PW_G);
  }
}

template <int MODE, int V>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_kernel_pwfwd(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ TcArgs a, const __grid_constant__ CUtensorMap tmW32, const PwArgs pw, const int n_pw) {
  mh_pdl_sync();
  extern __shared__ uint8_t smem_raw[];
  const int64_t pair = blockIdx.x >> 1, pairs = gridDim.x >> 1;
  if (pair < n_pw) pw_role(pw, &tmW32, (int)blockIdx.x, 2 * n_pw, smem_raw);
  else tc_body<MODE, V>(tmA, tmB, a, pair - n_pw, pairs - n_pw, smem_raw);
}

// ---- host side: tensor maps ------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  });
  return fn;
}

// 2-D row-major bf16 matrix [rows][cols]; box = [box_rows][64 cols] with the 128B swizzle.
int make_tmap(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { mh_set_error("cuTensorMapEncodeTiled entry point not found"); return MH_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { mh_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return MH_ERR_CUDA; }
  return MH_OK;
}

int num_sms() { return mh_num_sms(); }     // per device (capi.cu)

template <int MODE, int V>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const TcArgs& args, cudaStream_t st) {
  static MhDeviceOnce attr_once;               // one per template instance; kernel attributes are per device
  constexpr int smem = mode_smem_bytes(MODE);
  MH_CUDA_OK(mh_once_per_device(attr_once, [&] {
    return cudaFuncSetAttribute(tc_kernel<MODE, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  }));
  const int units = num_sms() / 2;
  // A-stationary kernels use a static schedule over exactly `units` pairs (pairs without work exit at once)
  int n = mode_astat(MODE) ? units : (int)std::min<int64_t>(args.total_tiles, units);
  // DW: the two d halves of a class tile (tiles 2k, 2k+1) run on neighbouring pairs in the same sweep
  if (MODE == MODE_DW && n > 1) n &= ~1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * n);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // see mh_launch (common.cuh)
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = mh_pdl_enabled() ? 2 : 1;
  MH_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_kernel<MODE, V>, ta, tb, args));
  return MH_OK;
}

int variant_of(const MhParams& p) {
  if (p.hard_kind == 1) return V_MV;
  if (p.hard_kind == 2) return V_CURR;
  if (p.scale_is_norm) return V_SPHERE;
  if (p.family == MH_ARCFACE) return V_PLAIN;
  return V_CLAMP;
}

template <int MODE>
int launch_variant(const CUtensorMap& ta, const CUtensorMap& tb, const TcArgs& args, cudaStream_t st) {
  switch (variant_of(args.p)) {
    case V_PLAIN: return launch<MODE, V_PLAIN>(ta, tb, args, st);
    case V_CLAMP: return launch<MODE, V_CLAMP>(ta, tb, args, st);
    case V_SPHERE: return launch<MODE, V_SPHERE>(ta, tb, args, st);
    case V_MV: return launch<MODE, V_MV>(ta, tb, args, st);
    default: return launch<MODE, V_CURR>(ta, tb, args, st);
  }
}

float family_umax(const MhParams& p) { return mh_family_umax(&p); }

// A-stationary schedule parameters for m_tiles <= units row tiles (see StatIter).
void make_sched(TcArgs& a, int units) {
  a.sG = units / a.m_tiles;
  a.sE = units - a.sG * a.m_tiles;
  const int64_t wfix = (int64_t)a.sG * a.m_tiles;
  a.n_fixed = a.sE == 0 ? a.n_tiles : (int)(((int64_t)a.n_tiles * wfix + (wfix + a.sE) / 2) / (wfix + a.sE));
}

int check_common(int64_t B, int64_t B_pad, int64_t C, int64_t C_pad) {
  MH_CHECK_ARG(B > 0 && B_pad >= B && B_pad % BMT == 0, "B_pad must be a multiple of 256 for the tensor-core path");
  MH_CHECK_ARG(C > 0 && C_pad >= C && C_pad % BN == 0, "C_pad must be a multiple of 256 for the tensor-core path");
  MH_CHECK_ARG(C_pad < (1ll << 31) && B_pad < (1ll << 31) && (C_pad / 128) * B_pad < (1ll << 31), "dimension too large");
  return MH_OK;
}

}  // namespace

// One statistics record per CTA pair and 128-column epilogue half: a pair accumulates its rows' statistics in registers
// over all the class tiles it sweeps.  (The argument is kept for ABI stability; the count no longer depends on C.)
extern "C" int64_t mh_fwd_num_tiles(int64_t C_pad) { (void)C_pad; return 2 * (int64_t)(num_sms() / 2); }

// records of rows a pair never visits stay at the merge identity (max = -inf, sums = 0)
__global__ void stats_identity_kernel(float* __restrict__ st, int64_t n_parts, int64_t B_pad, const int* gate, int gate_on) {
  mh_pdl_sync();
  if (gate_closed(gate, gate_on)) return;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = n_parts * MH_ST_PLANES * B_pad;
  if (i < n) st[i] = ((i / B_pad) % MH_ST_PLANES == MH_ST_M) ? -INFINITY : 0.f;
}

// Fixed-reference softmax (and with it the forward stash) is used when every possible non-target term
// exp2(z log2e - ref), ref = s log2e umax - 102, is a normal fp32 AND bf16 number with head-room for the row sums and
// for rho_i = s 2^(ref - lse): s log2e (umax + 1) <= 200, a fixed logit scale, and at least two classes per shard.
extern "C" int mh_tc_fixref_ok(const mh_config* cfg_host, int64_t C) {
  if (!cfg_host) return 0;
  const MhParams p = mh_make_params(cfg_host);
  if (p.scale_is_norm || C < 2 || !(p.s > 0.f)) return 0;
  return (p.s * MH_LOG2E * (family_umax(p) + 1.f) <= 200.f) ? 1 : 0;
}

// The forward stash additionally needs cos_ij to be recoverable from the stashed exponential (the backward rebuilds the
// projection term r_j from it): u = cos, or MV-Softmax's invertible u = w*cos + w - 1 on hard negatives.  CurricularFace's
// cos*(t + cos) with its cos-dependent derivative is not.
extern "C" int mh_tc_stash_ok(const mh_config* cfg_host, int64_t C) {
  if (!mh_tc_fixref_ok(cfg_host, C)) return 0;
  const MhParams p = mh_make_params(cfg_host);
  return (p.hard_kind == 0 || (p.hard_kind == 1 && p.hard_a > 1.f)) ? 1 : 0;
}

// Guarded stash: heads that fail the proof above (CurricularFace at s = 64: 277 binades; SphereFace: the scale is |x_i|;
// any family at s > 69) may still stash against the fixed reference ref_i = scale_i log2e umax - 102.  Nothing can
// overflow (u <= umax), but terms below 2^-126 flush to zero; that is harmless exactly when the row's sum is large
// against everything that can have been flushed, sum_j e_ij >= C 2^-102 (flushed mass <= C 2^-126: < 2^-24 relative, in
// the loss and in every gradient entry).  mh_finalize_rows_guarded checks this per row on the device; a batch with an
// unsafe row re-runs the general path (online-max forward, recomputed G) through launches gated on that flag.  Rows
// fail only when EVERY class logit sits > 100 nats below scale*umax, e.g. a handful of classes all anti-aligned with x.
extern "C" int mh_tc_stash_guarded_ok(const mh_config* cfg_host, int64_t C) {
  if (!cfg_host || C < 2) return 0;
  const MhParams p = mh_make_params(cfg_host);
  if (!p.scale_is_norm && !(p.s > 0.f)) return 0;
  return mh_tc_stash_ok(cfg_host, C) ? 0 : 1;
}

// Test hook (host only, no device work): the (pair, m_tile, n_tile) triples of the A-stationary schedule in
// execution order; returns the number of triples (<= cap are written) or a negative status.
extern "C" int64_t mh_tc_schedule_tiles(int units, int m_tiles, int n_tiles, int32_t* out, int64_t cap) {
  if (units <= 0 || m_tiles <= 0 || m_tiles > units || n_tiles <= 0) return MH_ERR_ARG;
  TcArgs a{};
  a.m_tiles = m_tiles; a.n_tiles = n_tiles;
  make_sched(a, units);
  int64_t cnt = 0;
  for (int p = 0; p < units; ++p) {
    StatIter it;
    it.init(a, p);
    int m, n;
    while (it.next(m, n)) {
      if (out && cnt < cap) { out[3 * cnt] = p; out[3 * cnt + 1] = m; out[3 * cnt + 2] = n; }
      ++cnt;
    }
  }
  return cnt;
}

// FWD / FWDS / BWD_G launch: A-stationary schedule; row tiles beyond `units` pairs go in further launches.
template <int MODE>
static int launch_s_tiles(const mh_config* cfg_host, const void* x_hat_bf16, int64_t B, int64_t B_pad, const void* w_hat_bf16,
                          int64_t C, int64_t C_pad, const float* rowp, int64_t ldp, const int32_t* label_local,
                          const float* state, const float* lse2, float* stats_tiles, void* bc_bf16, float* r_colsum,
                          cudaStream_t st, const int* gate = nullptr, int gate_on = 0) {
  const int units = num_sms() / 2;
  CUtensorMap tb;
  if (int e = make_tmap(&tb, w_hat_bf16, C_pad, MH_D, BN / 2)) return e;
  const int64_t rows_per_launch = (int64_t)units * BMT;
  if (stats_tiles) {
    const int64_t n = 2 * (int64_t)units * MH_ST_PLANES * B_pad;
    mh_launch(stats_identity_kernel, (unsigned)((n + 255) / 256), 256, 0, st, stats_tiles, 2 * units, B_pad, gate, gate_on);
  }
  for (int64_t r0 = 0; r0 < B_pad; r0 += rows_per_launch) {
    const int64_t rows = std::min(rows_per_launch, B_pad - r0);
    if (r0 >= B) break;                                               // only padding rows left
    CUtensorMap ta;
    if (int e = make_tmap(&ta, (const __nv_bfloat16*)x_hat_bf16 + r0 * MH_D, rows, MH_D, BM)) return e;
    TcArgs a{};
    a.m_tiles = (int)(rows / BMT); a.n_tiles = (int)(C_pad / BN); a.n_split = 1;
    a.k_blocks_total = MH_D / BK; a.k_blocks_per_split = a.k_blocks_total;
    a.total_tiles = (int64_t)a.m_tiles * a.n_tiles;
    make_sched(a, units);
    a.p = mh_make_params(cfg_host);
    a.fixref = mh_tc_fixref_ok(cfg_host, C);
    a.umax = family_umax(a.p);
    a.B = B - r0; a.C = C; a.B_pad = B_pad; a.C_pad = C_pad;
    a.rowp = rowp + r0; a.ldp = ldp; a.label_local = label_local + r0; a.state = state;
    a.lse2 = lse2 ? lse2 + r0 : nullptr;
    a.stats_tiles = stats_tiles ? stats_tiles + r0 : nullptr;
    a.G = bc_bf16 ? (__nv_bfloat16*)bc_bf16 + r0 * 128 : nullptr;
    a.rsum = r_colsum;
    a.gate = gate; a.gate_on = gate_on;
    if (int e = launch_variant<MODE>(ta, tb, a, st)) return e;
  }
  return MH_OK;
}

// stash_kind: 0 no stash, 1 the proven stash (mh_tc_stash_ok), 2 the guarded stash (mh_tc_stash_guarded_ok: the caller
// checks the row sums afterwards, see mh_step_forward).  gate: see TcArgs::gate.
extern "C" int mh_tc_forward_ex(const mh_config* cfg_host, const void* x_hat_bf16, int64_t B, int64_t B_pad, const void* w_hat_bf16,
                       int64_t C, int64_t C_pad, const float* rowp, int64_t ldp, const int32_t* label_local,
                       const float* state, float* stats_tiles, void* stash_bf16, int stash_kind, const int* gate,
                       int gate_on, void* stream) {
  MH_CHECK_ARG(cfg_host && x_hat_bf16 && w_hat_bf16 && rowp && label_local && state && stats_tiles, "null pointer");
  if (int e = check_common(B, B_pad, C, C_pad)) return e;
  MH_CHECK_ARG(ldp >= B_pad, "rowp pitch must cover B_pad");
  if (stash_bf16) {
    if (stash_kind == 2)
      MH_CHECK_ARG(mh_tc_stash_guarded_ok(cfg_host, C), "head not eligible for the guarded stash (see mh_tc_stash_guarded_ok)");
    else
      MH_CHECK_ARG(mh_tc_stash_ok(cfg_host, C), "head not eligible for the forward stash (see mh_tc_stash_ok)");
    return launch_s_tiles<MODE_FWDS>(cfg_host, x_hat_bf16, B, B_pad, w_hat_bf16, C, C_pad, rowp, ldp, label_local, state,
                                     nullptr, stats_tiles, stash_bf16, nullptr, (cudaStream_t)stream, gate, gate_on);
  }
  return launch_s_tiles<MODE_FWD>(cfg_host, x_hat_bf16, B, B_pad, w_hat_bf16, C, C_pad, rowp, ldp, label_local, state,
                                  nullptr, stats_tiles, nullptr, nullptr, (cudaStream_t)stream, gate, gate_on);
}

extern "C" int mh_tc_forward(const mh_config* cfg_host, const void* x_hat_bf16, int64_t B, int64_t B_pad,
                             const void* w_hat_bf16, int64_t C, int64_t C_pad, const float* rowp, int64_t ldp,
                             const int32_t* label_local, const float* state, float* stats_tiles, void* stash_bf16,
                             void* stream) {
  return mh_tc_forward_ex(cfg_host, x_hat_bf16, B, B_pad, w_hat_bf16, C, C_pad, rowp, ldp, label_local, state,
                            stats_tiles, stash_bf16, 1, nullptr, 0, stream);
}

extern "C" int mh_tc_backward_g_ex(const mh_config* cfg_host, const void* x_hat_bf16, int64_t B, int64_t B_pad,
                          const void* w_hat_bf16, int64_t C, int64_t C_pad, const float* rowp, int64_t ldp,
                          const int32_t* label_local, const float* state, const float* lse2, void* G_bf16,
                          float* r_colsum, const int* gate, int gate_on, void* stream) {
  MH_CHECK_ARG(cfg_host && x_hat_bf16 && w_hat_bf16 && rowp && label_local && state && lse2 && G_bf16, "null pointer");
  if (int e = check_common(B, B_pad, C, C_pad)) return e;
  MH_CHECK_ARG(ldp >= B_pad, "rowp pitch must cover B_pad");
  MH_CHECK_ARG(!(gate && r_colsum), "a gated backward-G launch takes no column sums");
  if (r_colsum) MH_CUDA_OK(cudaMemsetAsync(r_colsum, 0, sizeof(float) * C_pad, (cudaStream_t)stream));
  return launch_s_tiles<MODE_BWD_G>(cfg_host, x_hat_bf16, B, B_pad, w_hat_bf16, C, C_pad, rowp, ldp, label_local, state,
                                    lse2, nullptr, G_bf16, r_colsum, (cudaStream_t)stream, gate, gate_on);
}

extern "C" int mh_tc_backward_g(const mh_config* cfg_host, const void* x_hat_bf16, int64_t B, int64_t B_pad,
                                const void* w_hat_bf16, int64_t C, int64_t C_pad, const float* rowp, int64_t ldp,
                                const int32_t* label_local, const float* state, const float* lse2, void* G_bf16,
                                float* r_colsum, void* stream) {
  return mh_tc_backward_g_ex(cfg_host, x_hat_bf16, B, B_pad, w_hat_bf16, C, C_pad, rowp, ldp, label_local, state, lse2,
                               G_bf16, r_colsum, nullptr, 0, stream);
}

// side pass of the stash backward (NULL rho = plain dx GEMM)
struct DxSide {
  const float* rho = nullptr;
  float kappa = 0.f, inv_s2 = 0.f;
  int mv = 0;
  float ha = 1.f, hb = 0.f;
  const float* rowp = nullptr;
  int64_t ldp = 0;
  float* rsum = nullptr;
  int* sync_ws = nullptr;      // >= MH_DX_SYNC_INTS ints of scratch (NULL: no lockstep)
};

static int tc_backward_dx_impl(const void* G_bf16, int64_t B_pad, int64_t C_pad, const void* w_hat_bf16, float* out,
                               int* n_split_host, const DxSide& side, void* stream) {
  MH_CHECK_ARG(B_pad > 0 && B_pad % BMT == 0 && C_pad > 0 && C_pad % BN == 0, "bad padded shape");
  const int m_tiles = (int)(B_pad / BMT);
  const int kb_total = (int)(C_pad / BK);
  // Split K (the classes) so that the m_tiles x n_split tiles fill whole waves of the CTA pairs: cost = waves x (k-blocks
  // per split + the accumulator drain, which this single-buffered 256 x 512 tile does not overlap, ~8 k-blocks).
  // E.g. 8 GPUs (32 row tiles): 2 splits leave 10 of 74 pairs idle, 9 splits run 4 full waves (-10 %).  The fp32
  // partials are capped at 256 MB; a larger split count must pay by more than 3 % (measured: 37 instead of 18 splits
  // at 4 row tiles predicts -2 % and gains nothing, the combine kernel eats it).
  const int pairs = num_sms() / 2;
  const int n_cap = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(kb_total, 40), (256ll << 20) / (B_pad * MH_D * 4)));
  int n_split = 1;
  double best = 1e30;
  for (int n = 1; n <= n_cap; ++n) {
    const int per_n = (kb_total + n - 1) / n;
    const int n_eff = (kb_total + per_n - 1) / per_n;                 // no empty splits
    const int waves = (m_tiles * n_eff + pairs - 1) / pairs;
    const double cost = (double)waves * (per_n + 8);
    if (cost < best * 0.97) { best = cost; n_split = n_eff; }
  }
  static const int env_split = [] { const char* e = getenv("MH_DX_SPLIT"); return e ? atoi(e) : 0; }();   // experiments; read once
  if (env_split > 0) n_split = std::max(1, std::min(env_split, kb_total));
  int per = (kb_total + n_split - 1) / n_split;
  n_split = (kb_total + per - 1) / per;                 // no empty splits
  if (n_split_host) *n_split_host = n_split;
  if (!out) return MH_OK;
  MH_CHECK_ARG(G_bf16 && w_hat_bf16, "null pointer");
  CUtensorMap ta, tb;
  if (int e = make_tmap(&ta, G_bf16, (C_pad / 128) * B_pad, 128, BM)) return e;   // A = G (class-tiled), K-major
  if (int e = make_tmap(&tb, w_hat_bf16, C_pad, MH_D, 64)) return e;       // B = w^, MN-major boxes [64 k][64 d]
  TcArgs a{};
  a.m_tiles = m_tiles; a.n_tiles = 1; a.n_split = n_split;
  a.k_blocks_total = kb_total; a.k_blocks_per_split = per;
  a.total_tiles = (int64_t)m_tiles * n_split;
  a.B_pad = B_pad; a.C_pad = C_pad; a.B = B_pad; a.C = C_pad;
  a.out = out; a.out_split_stride = B_pad * MH_D;
  a.rho = side.rho; a.side_kappa = side.kappa; a.side_inv_s2 = side.inv_s2; a.rsum = side.rsum;
  a.side_mv = side.mv; a.side_ha = side.ha; a.side_hb = side.hb; a.rowp = side.rowp; a.ldp = side.ldp;
  // lockstep only when every tile is resident at once (one wave) and several row tiles share a split
  static const int env_lock = [] { const char* e = getenv("MH_DX_LOCKSTEP"); return e ? atoi(e) : 0; }();
  if (side.sync_ws && env_lock && m_tiles > 1 && a.total_tiles <= pairs && n_split <= MH_DX_SYNC_INTS) {
    MH_CUDA_OK(cudaMemsetAsync(side.sync_ws, 0, sizeof(int) * n_split, (cudaStream_t)stream));
    a.dx_sync = side.sync_ws;
  }
  return launch<MODE_DX, V_NONE>(ta, tb, a, (cudaStream_t)stream);
}

extern "C" int mh_tc_backward_dx(const void* G_bf16, int64_t B_pad, int64_t C_pad, const void* w_hat_bf16, float* out,
                                 int* n_split_host, int* sync_ws, void* stream) {
  DxSide side;
  side.sync_ws = sync_ws;
  return tc_backward_dx_impl(G_bf16, B_pad, C_pad, w_hat_bf16, out, n_split_host, side, stream);
}

extern "C" int mh_tc_backward_dx_stash(const mh_config* cfg_host, const void* stash_bf16, int64_t B_pad, int64_t C,
                                       int64_t C_pad, const void* w_hat_bf16, const float* rho, const float* rowp,
                                       int64_t ldp, float* out, float* r_colsum, int* n_split_host, int* sync_ws,
                                       void* stream) {
  MH_CHECK_ARG(cfg_host, "null pointer");
  MH_CHECK_ARG(!out || (rho && r_colsum && rowp && ldp >= B_pad), "null pointer");
  MH_CHECK_ARG(mh_tc_stash_ok(cfg_host, C), "head not eligible for the stash backward (see mh_tc_stash_ok)");
  const MhParams p = mh_make_params(cfg_host);
  const float s2 = p.s * MH_LOG2E;
  DxSide side;
  side.sync_ws = sync_ws;
  if (out) {
    side.rho = rho; side.kappa = (s2 * family_umax(p) - 102.f) / s2; side.inv_s2 = 1.f / s2; side.rsum = r_colsum;
    side.mv = (p.hard_kind == 1); side.ha = p.hard_a; side.hb = p.hard_b; side.rowp = rowp; side.ldp = ldp;
  }
  return tc_backward_dx_impl(stash_bf16, B_pad, C_pad, w_hat_bf16, out, n_split_host, side, stream);
}

static int launch_dw(const void* G_bf16, int64_t B_pad, int64_t C, int64_t C_pad, const void* x_bf16, TcArgs& a,
                     cudaStream_t st) {
  CUtensorMap ta, tb;
  if (int e = make_tmap(&ta, G_bf16, (C_pad / 128) * B_pad, 128, 64)) return e;    // A = G^T (class-tiled G), MN-major boxes
  if (int e = make_tmap(&tb, x_bf16, B_pad, MH_D, 64)) return e;                   // B = x^, MN-major boxes [64 rows][64 d]
  a.m_tiles = (int)(C_pad / BMT); a.n_tiles = 2; a.n_split = 1;
  a.k_blocks_total = (int)(B_pad / BK); a.k_blocks_per_split = a.k_blocks_total;
  a.total_tiles = (int64_t)a.m_tiles * 2;
  a.B_pad = B_pad; a.C_pad = C_pad; a.B = B_pad; a.C = C;
  a.out_split_stride = 0;
  return launch<MODE_DW, V_NONE>(ta, tb, a, st);
}

extern "C" int mh_tc_backward_dw(const void* G_bf16, int64_t B_pad, int64_t C_pad, const void* x_hat_bf16,
                                 float* dw_hat, void* stream) {
  MH_CHECK_ARG(G_bf16 && x_hat_bf16 && dw_hat, "null pointer");
  MH_CHECK_ARG(B_pad > 0 && B_pad % BM == 0 && C_pad > 0 && C_pad % BN == 0, "bad padded shape");
  MH_CHECK_ARG(((uintptr_t)dw_hat & 15) == 0, "dw_hat must be 16-byte aligned");
  TcArgs a{};
  a.out = dw_hat; a.raw_dw = 1; a.layout = MH_LAYOUT_CD; a.ld = MH_D;
  a.w_hat = (const __nv_bfloat16*)x_hat_bf16;            // never read in raw mode; any valid pointer
  return launch_dw(G_bf16, B_pad, C_pad, C_pad, x_hat_bf16, a, (cudaStream_t)stream);
}

extern "C" int mh_tc_backward_dw_fused(const void* G_bf16, int64_t B_pad, int64_t C, int64_t C_pad, const void* x_hat_bf16,
                                       const void* w_hat_bf16, const float* inv_norm, const float* r_colsum,
                                       int r_parts, const float* gscal, int layout, float* dW, int64_t ld, void* stream) {
  MH_CHECK_ARG(G_bf16 && x_hat_bf16 && w_hat_bf16 && inv_norm && r_colsum && gscal && dW, "null pointer");
  MH_CHECK_ARG(B_pad > 0 && B_pad % BM == 0 && C_pad > 0 && C_pad % BN == 0 && C > 0 && C <= C_pad, "bad padded shape");
  MH_CHECK_ARG(layout == MH_LAYOUT_CD || layout == MH_LAYOUT_DC, "unknown layout");
  MH_CHECK_ARG(layout != MH_LAYOUT_CD || (ld % 4 == 0 && ((uintptr_t)dW & 15) == 0), "CD dW must be 16-byte aligned");
  TcArgs a{};
  a.out = dW; a.raw_dw = 0; a.layout = layout; a.ld = ld;
  a.w_hat = (const __nv_bfloat16*)w_hat_bf16; a.inv_norm = inv_norm; a.gscal = gscal;
  a.rsum = const_cast<float*>(r_colsum); a.rsum_parts = r_parts;
  MH_CHECK_ARG(r_parts >= 1, "r_parts must be >= 1");
  if (int e = make_tmap(&a.tmW, w_hat_bf16, C_pad, MH_D, BM)) return e;       // epilogue operand: boxes [64 d][128 classes]
  return launch_dw(G_bf16, B_pad, C, C_pad, x_hat_bf16, a, (cudaStream_t)stream);
}

// dW with the projection taken from the accumulators (no r_colsum input): rpart_ws [4 * C_pad] floats, flag_ws
// [C_pad / 128] ints, both scratch owned by the caller; flag_ws is zeroed here (stream-ordered).
extern "C" int mh_tc_backward_dw_proj(const void* G_bf16, int64_t B_pad, int64_t C, int64_t C_pad, const void* x_hat_bf16,
                                      const void* w_hat_bf16, const float* inv_norm, const float* gscal, int layout,
                                      float* dW, int64_t ld, float* rpart_ws, int* flag_ws, void* stream) {
  MH_CHECK_ARG(G_bf16 && x_hat_bf16 && w_hat_bf16 && inv_norm && gscal && dW && rpart_ws && flag_ws, "null pointer");
  MH_CHECK_ARG(B_pad > 0 && B_pad % BM == 0 && C_pad > 0 && C_pad % BN == 0 && C > 0 && C <= C_pad, "bad padded shape");
  MH_CHECK_ARG(layout == MH_LAYOUT_CD || layout == MH_LAYOUT_DC, "unknown layout");
  MH_CHECK_ARG(layout != MH_LAYOUT_CD || (ld % 4 == 0 && ((uintptr_t)dW & 15) == 0), "CD dW must be 16-byte aligned");
  MH_CUDA_OK(cudaMemsetAsync(flag_ws, 0, sizeof(int) * (size_t)(C_pad / BM), (cudaStream_t)stream));
  TcArgs a{};
  a.out = dW; a.raw_dw = 0; a.layout = layout; a.ld = ld;
  a.w_hat = (const __nv_bfloat16*)w_hat_bf16; a.inv_norm = inv_norm; a.gscal = gscal;
  a.rsum = nullptr; a.rsum_parts = 0; a.rpart = rpart_ws; a.rflag = flag_ws;
  if (int e = make_tmap(&a.tmW, w_hat_bf16, C_pad, MH_D, BM)) return e;
  return launch_dw(G_bf16, B_pad, C, C_pad, x_hat_bf16, a, (cudaStream_t)stream);
}

// ---- merged backward: dx^ partials and dW in ONE persistent kernel (see tc_kernel_dxdw) ---------------------------------
// Split of the CTA pairs between the two roles: the dx GEMM gets m_tiles * n_split pairs (~38 % of the chip: the two GEMMs
// have the same FLOPs, the dW role also writes 4.1 GB of fp32 dW and projects), the dW role an even number of the rest
// (partner exchange).
static int dxdw_split(int m_tiles, int pairs) {
  if (m_tiles < 1 || m_tiles > pairs / 2) return 0;
  static const double frac = [] { const char* e = getenv("MH_DXDW_FRAC"); return e ? atof(e) : 0.38; }();   // measured best of 0.38 / 0.44 / 0.49
  int n_split = std::max(1, (int)(frac * pairs / m_tiles + 0.5));
  while (n_split >= 1) {
    const int n_dw = pairs - m_tiles * n_split;
    if (n_dw >= 2 && n_dw % 2 == 0) return n_split;
    --n_split;
  }
  return 0;
}

extern "C" int mh_tc_backward_dxdw(const void* G_bf16, int64_t B_pad, int64_t C, int64_t C_pad, const void* w_hat_bf16,
                                   const void* xs_bf16, const float* inv_norm, const float* gscal, int layout, float* dW,
                                   int64_t ld, float* dxhat_part, int* n_split_host, float* rpart_ws, int* flag_ws,
                                   int* prog_ws, void* stream) {
  MH_CHECK_ARG(B_pad > 0 && B_pad % BMT == 0 && C_pad > 0 && C_pad % BN == 0 && C > 0 && C <= C_pad, "bad padded shape");
  const int pairs = num_sms() / 2;
  const int m_tiles = (int)(B_pad / BMT);
  const int kb_total = (int)(C_pad / BK);
  // worth it only when every pair streams many class tiles; otherwise the caller runs the two kernels back to back
  int n_split = (C_pad / BMT >= 8 * (int64_t)pairs) ? dxdw_split(m_tiles, pairs) : 0;
  if (n_split > 0 && (int64_t)n_split * B_pad * MH_D * 4 > (256ll << 20)) n_split = 0;
  if (n_split_host) *n_split_host = n_split;
  if (!dxhat_part) return MH_OK;                                       // query
  MH_CHECK_ARG(n_split > 0, "shape not eligible for the merged backward (query with dxhat_part == NULL first)");
  MH_CHECK_ARG(G_bf16 && w_hat_bf16 && xs_bf16 && inv_norm && gscal && dW && rpart_ws && flag_ws && prog_ws, "null pointer");
  MH_CHECK_ARG(layout == MH_LAYOUT_CD || layout == MH_LAYOUT_DC, "unknown layout");
  MH_CHECK_ARG(layout != MH_LAYOUT_CD || (ld % 4 == 0 && ((uintptr_t)dW & 15) == 0), "CD dW must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  MH_CUDA_OK(cudaMemsetAsync(flag_ws, 0, sizeof(int) * (size_t)(C_pad / BM), st));
  MH_CUDA_OK(cudaMemsetAsync(prog_ws, 0, sizeof(int) * 2, st));
  // DX role
  CUtensorMap ta_dx, tb_dx, ta_dw, tb_dw;
  if (int e = make_tmap(&ta_dx, G_bf16, (C_pad / 128) * B_pad, 128, BM)) return e;
  if (int e = make_tmap(&tb_dx, w_hat_bf16, C_pad, MH_D, 64)) return e;
  TcArgs ax{};
  ax.m_tiles = m_tiles; ax.n_tiles = 1; ax.n_split = n_split;
  ax.k_blocks_total = kb_total; ax.k_blocks_per_split = (kb_total + n_split - 1) / n_split;
  ax.total_tiles = (int64_t)m_tiles * n_split;
  ax.B_pad = B_pad; ax.C_pad = C_pad; ax.B = B_pad; ax.C = C_pad;
  ax.out = dxhat_part; ax.out_split_stride = B_pad * MH_D;
  ax.dx_chunk = BMT / BK;                                              // 4 k-blocks = one 256-class tile of the dW role
  ax.prog = prog_ws;
  static const int env_ahead = [] { const char* e = getenv("MH_PROG_AHEAD"); return e ? atoi(e) : 0; }();   // experiments
  ax.prog_ahead = env_ahead > 0 ? env_ahead : PROG_AHEAD;
  // DW role (self-projecting)
  if (int e = make_tmap(&ta_dw, G_bf16, (C_pad / 128) * B_pad, 128, 64)) return e;
  if (int e = make_tmap(&tb_dw, xs_bf16, B_pad, MH_D, 64)) return e;
  TcArgs aw{};
  aw.out = dW; aw.raw_dw = 0; aw.layout = layout; aw.ld = ld;
  aw.w_hat = (const __nv_bfloat16*)w_hat_bf16; aw.inv_norm = inv_norm; aw.gscal = gscal;
  aw.rpart = rpart_ws; aw.rflag = flag_ws; aw.prog = prog_ws; aw.prog_ahead = ax.prog_ahead;
  if (int e = make_tmap(&aw.tmW, w_hat_bf16, C_pad, MH_D, BM)) return e;
  aw.m_tiles = (int)(C_pad / BMT); aw.n_tiles = 2; aw.n_split = 1;
  aw.k_blocks_total = (int)(B_pad / BK); aw.k_blocks_per_split = aw.k_blocks_total;
  aw.total_tiles = (int64_t)aw.m_tiles * 2;
  aw.B_pad = B_pad; aw.C_pad = C_pad; aw.B = B_pad; aw.C = C;
  static MhDeviceOnce attr_once;
  constexpr int smem = mode_smem_bytes(MODE_DW) > mode_smem_bytes(MODE_DX) ? mode_smem_bytes(MODE_DW) : mode_smem_bytes(MODE_DX);
  MH_CUDA_OK(mh_once_per_device(attr_once, [&] {
    return cudaFuncSetAttribute(tc_kernel_dxdw, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  }));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // see mh_launch (common.cuh)
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = mh_pdl_enabled() ? 2 : 1;
  const int n_dx = m_tiles * n_split;
  MH_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_kernel_dxdw, ta_dx, tb_dx, ax, ta_dw, tb_dw, aw, n_dx));
  return MH_OK;
}

// ---- merged W prologue + forward: host side ------------------------------------------------------------------------
// fp32 [rows][512] row-major (pitch ld floats), boxes of [PW_G rows][256 columns], no swizzle, zero fill out of bounds
static int make_tmap_w32(CUtensorMap* out, const float* base, int64_t rows, int64_t ld) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { mh_set_error("cuTensorMapEncodeTiled entry point not found"); return MH_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)MH_D, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {256u, (cuuint32_t)PW_G};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { mh_set_error("cuTensorMapEncodeTiled (fp32 W) failed with CUresult %d", (int)r); return MH_ERR_CUDA; }
  return MH_OK;
}

template <int MODE, int V>
static int launch_pwfwd(const CUtensorMap& ta, const CUtensorMap& tb, const TcArgs& args, const CUtensorMap& tw32,
                        const PwArgs& pw, int n_pw, int pairs, cudaStream_t st) {
  static MhDeviceOnce attr_once;
  constexpr int smem = mode_smem_bytes(MODE);
  MH_CUDA_OK(mh_once_per_device(attr_once, [&] {
    return cudaFuncSetAttribute(tc_kernel_pwfwd<MODE, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  }));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // see mh_launch (common.cuh)
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = mh_pdl_enabled() ? 2 : 1;
  MH_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_kernel_pwfwd<MODE, V>, ta, tb, args, tw32, pw, n_pw));
  return MH_OK;
}

// pairs given to the prologue role: the forward keeps a multiple of m_tiles pairs (no left-over pairs, whose tail tiles
// would have to wait for the whole prologue), the prologue role gets the rest (>= MH_PW_PAIRS, default 12)
static int pwfwd_split(int m_tiles, int pairs) {
  static const int want = [] { const char* e = getenv("MH_PW_PAIRS"); return e ? atoi(e) : 12; }();
  if (m_tiles < 1 || want < 2) return 0;
  const int n_fwd = ((pairs - want) / m_tiles) * m_tiles;
  const int n_pw = pairs - n_fwd;
  return (n_fwd >= m_tiles && n_pw <= pairs / 3) ? n_pw : 0;
}

extern "C" int mh_tc_forward_pw(const mh_config* cfg_host, const void* x_hat_bf16, int64_t B, int64_t B_pad, const float* W,
                                int layout, int64_t ld, void* w_hat_bf16, float* inv_norm, int64_t C, int64_t C_pad,
                                const float* rowp, int64_t ldp, const int32_t* label_local, const float* state,
                                float* stats_tiles, void* stash_bf16, int* ready_ws, int* eligible_host, void* stream) {
  MH_CHECK_ARG(cfg_host, "null pointer");
  if (int e = check_common(B, B_pad, C, C_pad)) return e;
  const int pairs = num_sms() / 2;
  const int m_tiles = (int)(B_pad / BMT);
  const MhParams p = mh_make_params(cfg_host);
  const int v = variant_of(p);
  int n_pw = 0;
  // [C, 512] parameters only (ArcFace, SphereFace, MV-Softmax), one launch of row tiles, many class tiles per pair
  if (layout == MH_LAYOUT_CD && ld % 4 == 0 && p.family != MH_VPL_ARC && (v == V_PLAIN || v == V_MV || (v == V_SPHERE && !stash_bf16)) &&
      C_pad / BN >= 8 * (int64_t)pairs)
    n_pw = pwfwd_split(m_tiles, pairs);
  if (eligible_host) *eligible_host = n_pw > 0 ? 1 : 0;
  if (!x_hat_bf16) return MH_OK;                                           // query
  MH_CHECK_ARG(n_pw > 0, "shape / head not eligible for the merged prologue + forward (query with x_hat == NULL first)");
  MH_CHECK_ARG(W && w_hat_bf16 && inv_norm && rowp && label_local && state && stats_tiles && ready_ws, "null pointer");
  MH_CHECK_ARG(((uintptr_t)W & 15) == 0 && ldp >= B_pad, "W must be 16-byte aligned; rowp pitch must cover B_pad");
  if (stash_bf16) MH_CHECK_ARG(mh_tc_stash_ok(cfg_host, C), "head not eligible for the forward stash (see mh_tc_stash_ok)");
  cudaStream_t st = (cudaStream_t)stream;
  const int n_fwd = pairs - n_pw;
  const int n_ct = (int)(C_pad / BN);
  MH_CUDA_OK(cudaMemsetAsync(ready_ws, 0, sizeof(int) * (size_t)(n_ct + 1), st));
  {
    const int64_t n = 2 * (int64_t)pairs * MH_ST_PLANES * B_pad;
    mh_launch(stats_identity_kernel, (unsigned)((n + 255) / 256), 256, 0, st, stats_tiles, 2 * pairs, B_pad, nullptr, 0);
  }
  CUtensorMap ta, tb;
  if (int e = make_tmap(&tb, w_hat_bf16, C_pad, MH_D, BN / 2)) return e;
  if (int e = make_tmap(&ta, x_hat_bf16, B_pad, MH_D, BM)) return e;
  TcArgs a{};
  a.m_tiles = m_tiles; a.n_tiles = n_ct; a.n_split = 1;
  a.k_blocks_total = MH_D / BK; a.k_blocks_per_split = a.k_blocks_total;
  a.total_tiles = (int64_t)a.m_tiles * a.n_tiles;
  make_sched(a, n_fwd);
  a.p = p;
  a.fixref = mh_tc_fixref_ok(cfg_host, C);
  a.umax = family_umax(a.p);
  a.B = B; a.C = C; a.B_pad = B_pad; a.C_pad = C_pad;
  a.rowp = rowp; a.ldp = ldp; a.label_local = label_local; a.state = state;
  a.stats_tiles = stats_tiles;
  a.G = (__nv_bfloat16*)stash_bf16;
  a.w_ready = ready_ws; a.fwd_front = ready_ws + n_ct;
  PwArgs pw{};
  pw.W = W; pw.C = C; pw.C_pad = C_pad; pw.ld = ld;
  pw.what = (__nv_bfloat16*)w_hat_bf16; pw.inv_norm = inv_norm; pw.ready = ready_ws; pw.fwd_front = ready_ws + n_ct;
  CUtensorMap tw32;
  if (int e = make_tmap_w32(&tw32, W, C, ld)) return e;
  if (stash_bf16) {
    if (v == V_PLAIN) return launch_pwfwd<MODE_FWDS, V_PLAIN>(ta, tb, a, tw32, pw, n_pw, pairs, st);
    return launch_pwfwd<MODE_FWDS, V_MV>(ta, tb, a, tw32, pw, n_pw, pairs, st);
  }
  if (v == V_PLAIN) return launch_pwfwd<MODE_FWD, V_PLAIN>(ta, tb, a, tw32, pw, n_pw, pairs, st);
  if (v == V_MV) return launch_pwfwd<MODE_FWD, V_MV>(ta, tb, a, tw32, pw, n_pw, pairs, st);
  return launch_pwfwd<MODE_FWD, V_SPHERE>(ta, tb, a, tw32, pw, n_pw, pairs, st);
}
