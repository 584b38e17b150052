// Tensor-core path of the margin head for sm_100a: tcgen05.mma with TMEM accumulators, TMA-fed
// 128B-swizzled shared-memory stages, mbarrier pipelines, warp-specialised persistent CTAs.
//
// One kernel skeleton, four modes (template parameter):
//   FWD    S = x^ w^T tile  -> clamp/margin/scale -> online max/sum-exp + rank count per row.
//          Nothing of size B x C is written (replaces criterion.py:267-301 & siblings +
//          nn.CrossEntropyLoss, model_utils.py:179, + accuracy, metrics.py:3-16).
//   BWD_G  recompute the S tile -> G = (P - Y) * dz/dcos as bf16, class-tiled [C_pad/128][B_pad][128].
//          Also accumulates r_j = sum_i G_ij * cos_ij (= w^_j . dw^_j), the normalise-backward projection.
//   DX     dx^ partials = G . w^      (A = G K-major,  B = w^ MN-major, split over classes; one 128 x 512
//          accumulator = all 512 TMEM columns per CTA, so every G byte is read by exactly one CTA)
//   DW     dw^          = G^T . x^    (A = G MN-major, B = x^ MN-major), raw fp32 output
//   DWF    same GEMM, epilogue writes dW_j = g (dw^_j - w^_j r_j) / |w_j| straight into the parameter layout
//
// CTA = 384 threads: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-11
// epilogue (thread = one accumulator row x one 128-column half).  Tile 128 x 256 x 64, 4 smem stages
// (48 KB each), two 256-column TMEM accumulators so the epilogue of tile t overlaps the MMAs of t+1.
#include "common.cuh"
#include <cuda.h>
#include <cstdlib>
#include <mutex>
#include <map>
#include <tuple>

namespace {

constexpr int BM = 128, BN = 256, BK = 64, MAX_STAGES = 6;
constexpr int A_STAGE_BYTES = BM * BK * 2;      // 16 KB
constexpr int NUM_EPI_WARPS = 8;
constexpr int EPI_WARP0 = 4;
constexpr int NUM_THREADS = 32 * (EPI_WARP0 + NUM_EPI_WARPS);
constexpr uint32_t TMEM_COLS = 512;

enum { MODE_FWD = 0, MODE_BWD_G = 1, MODE_DX = 2, MODE_DW = 3, MODE_DWF = 4 };
// Per-mode pipeline configuration.  cta2 = the cta_group::2 variant: a cluster of two CTAs (one SM pair) computes a
// 256-row tile; each CTA keeps its own 128 A rows + 128 accumulator lanes and loads only HALF of the B tile, the
// tensor cores of both SMs read both halves.  Per-CTA bytes per MMA drop by a third to a half (the 1-CTA kernels
// were bound by per-SM TMA/L2 request throughput), which also buys deeper pipelines.
//
// A-stationary (cta2 FWD / BWD_G): the pair's x^ tile (128 rows x 512 per CTA = 128 KB) stays resident in shared
// memory while the pair sweeps class tiles, so only w^ streams (16 KB per CTA per k-block): L2 -> SM traffic per
// tile halves (64 -> 32 B/clk/SM; the non-stationary kernels sat on the ~6.3 KB/clk chip-wide L2 delivery cap).
constexpr bool mode_astat(int mode, bool cta2) { return cta2 && (mode == MODE_FWD || mode == MODE_BWD_G); }
constexpr int A_RESIDENT_BYTES = BM * MH_D * 2;                                   // 128 KB
// Output staging for coalesced global stores, per epilogue warp:
//   BWD_G: 32 rows x 128 B (64 bf16 columns), XOR-swizzled 16 B pieces, flushed every two 32-column chunks;
//   DW/DWF: 32 rows x 256 B (+16 B pad per row), fp32 dw^ rows.
constexpr int STG_ROW_BYTES = 256 + 16;
constexpr int mode_stg_warp_bytes(int mode) {
  return mode == MODE_BWD_G ? 32 * 128 : ((mode == MODE_DW || mode == MODE_DWF) ? 32 * STG_ROW_BYTES : 0);
}
constexpr int mode_bn(int mode) { return mode == MODE_DX ? 512 : 256; }          // accumulator columns per tile
constexpr int mode_nbuf(int mode) { return mode == MODE_DX ? 1 : 2; }            // TMEM accumulators in flight
constexpr int mode_stage_bytes(int mode, bool cta2) {
  return (mode_astat(mode, cta2) ? 0 : A_STAGE_BYTES) + (mode_bn(mode) / (cta2 ? 2 : 1)) * BK * 2;
}
constexpr int mode_stages(int mode, bool cta2) {
  if (mode_astat(mode, cta2)) return mode == MODE_FWD ? 6 : 4;
  if (cta2) return 4;
  return mode == MODE_DX ? 2 : (mode_stg_warp_bytes(mode) ? 3 : 4);
}
constexpr int mode_smem_bytes(int mode, bool cta2) {
  return (mode_astat(mode, cta2) ? A_RESIDENT_BYTES : 0) + mode_stages(mode, cta2) * mode_stage_bytes(mode, cta2) +
         1024 /*align slack*/ + 256 /*barriers*/ + NUM_EPI_WARPS * mode_stg_warp_bytes(mode);
}
static_assert(mode_smem_bytes(MODE_BWD_G, false) <= 232448 && mode_smem_bytes(MODE_FWD, false) <= 232448 &&
              mode_smem_bytes(MODE_DX, false) <= 232448 && mode_smem_bytes(MODE_BWD_G, true) <= 232448 &&
              mode_smem_bytes(MODE_FWD, true) <= 232448 && mode_smem_bytes(MODE_DX, true) <= 232448 &&
              mode_smem_bytes(MODE_DWF, true) <= 232448 && mode_smem_bytes(MODE_DWF, false) <= 232448, "smem budget");

struct TcArgs {
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
  __nv_bfloat16* G;
  float* out;
  int64_t out_split_stride;
  float* rsum;                 // BWD_G: r_j accumulation target [C_pad] (zeroed by the caller); DWF: read
  const __nv_bfloat16* w_hat;  // DWF
  const float* inv_norm;       // DWF
  const float* gscal;          // DWF
  int layout;                  // DWF: parameter layout of dW
  int64_t ld;                  // DWF: row pitch of dW
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
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
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// L2 prefetch of a TMA box (no shared memory, no barrier): hides HBM latency beyond the smem pipeline depth.
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
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
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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
  int split;         // DX split index
  int n_tile;        // FWD: class-tile index
  int m_tile;        // A-stationary modes: row-tile index (the resident x^ tile)
};

// Tile -> work.  BMT = rows of the (pair) tile: 128, or 256 with cta_group::2; `rank` selects this CTA's 128 rows.
template <int MODE, bool CTA2>
__device__ __forceinline__ Work get_work(const TcArgs& a, int64_t t, int rank) {
  constexpr int BMT = CTA2 ? 2 * BM : BM;
  Work w;
  w.split = 0;
  w.n_tile = 0;
  w.m_tile = 0;
  if (MODE == MODE_FWD || MODE == MODE_BWD_G) {
    int m = (int)(t % a.m_tiles), n = (int)(t / a.m_tiles);
    w.m0 = m * BMT + rank * BM; w.n0 = n * BN; w.kb0 = 0; w.kb1 = MH_D / BK; w.n_tile = n;
  } else if (MODE == MODE_DX) {
    w.split = (int)(t / a.m_tiles);
    w.m0 = (int)(t % a.m_tiles) * BMT + rank * BM;
    w.n0 = 0;
    w.kb0 = w.split * a.k_blocks_per_split;
    w.kb1 = min(a.k_blocks_total, w.kb0 + a.k_blocks_per_split);
  } else {
    w.m0 = (int)(t >> 1) * BMT + rank * BM;
    w.n0 = (int)(t & 1) * BN;
    w.kb0 = 0; w.kb1 = a.k_blocks_total;
  }
  return w;
}

// A-stationary tile schedule (cta2 FWD / BWD_G).  `units` CTA pairs, m_tiles <= units row tiles of 256 rows:
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

// The tile sequence of one CTA (pair); the producer, MMA and epilogue roles all walk the same sequence.
template <int MODE, bool CTA2>
struct TileLoop {
  static constexpr bool AS = mode_astat(MODE, CTA2);
  StatIter si;
  int64_t t, npid;
  __device__ __forceinline__ void init(const TcArgs& a, int64_t pid, int64_t npid_) {
    t = pid; npid = npid_;
    if (AS) si.init(a, (int)pid);
  }
  __device__ __forceinline__ bool next(const TcArgs& a, int rank, Work& w) {
    if (AS) {
      int m, n;
      if (!si.next(m, n)) return false;
      w.m0 = m * 2 * BM + rank * BM; w.n0 = n * BN; w.kb0 = 0; w.kb1 = MH_D / BK;
      w.split = 0; w.n_tile = n; w.m_tile = m;
      return true;
    }
    if (t >= a.total_tiles) return false;
    w = get_work<MODE, CTA2>(a, t, rank);
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
  int tcol;          // tile-local target column, or -1
  bool valid;        // row < B
};

struct FwdAcc {
  float m, l, ez;
  int cnt;
};

template <int V>
__device__ __forceinline__ float elem_u(float raw, float lo, float hi, float thr, float ha, float hb, float& c_out) {
  float c = raw;
  if (V != V_PLAIN) c = fminf(fmaxf(raw, lo), hi);
  c_out = c;
  if (V == V_MV) return (c > thr) ? fmaf(ha, c, hb) : c;
  if (V == V_CURR) return (c > thr) ? c * (ha + c) : c;
  return c;
}
template <int V>
__device__ __forceinline__ float elem_du(float raw, float c, float thr, float ha) {
  float du = 1.f;
  if (V == V_MV) du = (c > thr) ? ha : 1.f;
  if (V == V_CURR) du = (c > thr) ? (ha + 2.f * c) : 1.f;
  if (V != V_PLAIN && c != raw) du = 0.f;            // clamp passes gradient only inside [lo, hi]
  return du;
}

// One 32-column chunk of the forward: online max / sum-exp (log2 domain), rank count, optional sum e*u.
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
      __nv_bfloat162 b = __floats2bfloat162_rn(g2[0], g2[1]);
      pk[k >> 1] = *reinterpret_cast<uint32_t*>(&b);
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
      __nv_bfloat162 b = __floats2bfloat162_rn(g2[0], g2[1]);
      pk[k >> 1] = *reinterpret_cast<uint32_t*>(&b);
    }
  }
}

template <int MODE, int V, bool CTA2>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
  constexpr int STAGES = mode_stages(MODE, CTA2);
  constexpr int STAGE_BYTES = mode_stage_bytes(MODE, CTA2);
  constexpr int NCTA = CTA2 ? 2 : 1;
  constexpr int GW = BN / NCTA;                      // B columns (of one 256-wide UMMA group) held by this CTA
  const int rank = CTA2 ? (int)cluster_ctarank() : 0;
  const int64_t pid = CTA2 ? (blockIdx.x >> 1) : blockIdx.x;        // tile-scheduling unit: CTA or CTA pair
  const int64_t npid = CTA2 ? (gridDim.x >> 1) : gridDim.x;
  constexpr int BNT = mode_bn(MODE);                 // accumulator columns of one tile (256, DX: 512)
  constexpr int NBUF = mode_nbuf(MODE);
  constexpr bool IS_DW = (MODE == MODE_DW || MODE == MODE_DWF);
  constexpr bool AS = mode_astat(MODE, CTA2);        // x^ tile resident in smem, only w^ streams
  constexpr int A_IN_STAGE = AS ? 0 : A_STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;           // SWIZZLE_128B needs 1024 B alignment
  const uint32_t ares_base = smem_base;                             // AS: resident A, 8 k-blocks x 16 KB
  const uint32_t tiles_base = smem_base + (AS ? A_RESIDENT_BYTES : 0);
  uint8_t* tiles_ptr = smem_raw + (tiles_base - raw_addr);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles_ptr + STAGES * STAGE_BYTES);
  uint8_t* stg_all = tiles_ptr + STAGES * STAGE_BYTES + 256;        // output staging (BWD_G / DW / DWF only)
  const uint32_t bar_full = smem_u32(bars);                         // [STAGES]
  const uint32_t bar_empty = bar_full + 8 * MAX_STAGES;             // [STAGES]
  const uint32_t bar_tfull = bar_empty + 8 * MAX_STAGES;            // [2]
  const uint32_t bar_tempty = bar_tfull + 16;                       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);
  const uint32_t bar_afull = bar_full + 8 * (2 * MAX_STAGES + 5);   // AS: resident A landed
  const uint32_t bar_afree = bar_afull + 8;                         // AS: every MMA that read the resident A retired

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, NCTA);                // pair: leader's expect_tx arrive + the peer producer's arrive
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, NUM_EPI_WARPS * NCTA);   // one arrive per epilogue warp (of both CTAs)
    }
    mbar_init(bar_afull, NCTA);
    mbar_init(bar_afree, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CTA2) tmem_alloc_2sm(smem_u32(tmem_slot), TMEM_COLS); else tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  }
  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  constexpr bool A_MN = IS_DW;
  constexpr bool B_MN = (MODE == MODE_DX || IS_DW);

  if (warp == 0) {
    // =============================== TMA producer (both CTAs of a pair) ===============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      TileLoop<MODE, CTA2> tl;
      tl.init(a, pid, npid);
      Work w;
      int res_m = -1;
      uint32_t tile_j = 0;
      while (tl.next(a, rank, w)) {
        if (AS && w.m_tile != res_m) {
          // (re)load the resident x^ tile: wait until every MMA of the previous tiles has retired
          if (tile_j > 0) mbar_wait(bar_afree, (tile_j - 1) & 1);
          if (rank == 0) mbar_expect_tx(bar_afull, NCTA * A_RESIDENT_BYTES); else mbar_arrive_cluster(bar_afull, 0);
          for (int kb = 0; kb < MH_D / BK; ++kb) {
            if (CTA2) tma_load_2d_2sm(ares_base + kb * A_STAGE_BYTES, &tmA, bar_afull, kb * BK, w.m0);
            else tma_load_2d(ares_base + kb * A_STAGE_BYTES, &tmA, bar_afull, kb * BK, w.m0);
          }
          res_m = w.m_tile;
        }
        ++tile_j;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t sa = tiles_base + stage * STAGE_BYTES;
          const uint32_t sb = sa + A_IN_STAGE;
          const uint32_t fb = bar_full + 8 * stage;
          if (!CTA2) {
            mbar_expect_tx(fb, STAGE_BYTES);
          } else if (rank == 0) {
            mbar_expect_tx(fb, 2 * STAGE_BYTES);                            // bytes of both CTAs land on the leader's barrier
          } else {
            mbar_arrive_cluster(fb, 0);
          }
          auto load = [&](uint32_t dst, const CUtensorMap* m, int c0, int c1) {
            if (CTA2) tma_load_2d_2sm(dst, m, fb, c0, c1); else tma_load_2d(dst, m, fb, c0, c1);
          };
          if (MODE == MODE_DX) {
            // A = G, class-tiled [C_pad/128][B_pad][128]: k-block kb = classes 64kb.. -> slab kb/2, columns (kb&1)*64
            load(sa, &tmA, (kb & 1) * 64, (kb >> 1) * (int)a.B_pad + w.m0);  // box [64 k][128 rows]
          } else if (AS) {
            // x^ is resident
          } else if (!A_MN) {
            load(sa, &tmA, kb * BK, w.m0);                                  // box [64 k][128 rows]
          } else {
            // A = G^T from the class-tiled G: this CTA's 128 classes are slab m0/128; box [64 classes][64 rows]
#pragma unroll
            for (int bx = 0; bx < BM / 64; ++bx)
              load(sa + bx * 8192, &tmA, 64 * bx, (w.m0 / 128) * (int)a.B_pad + kb * BK);
          }
          if (!B_MN) {
            load(sb, &tmB, kb * BK, w.n0 + rank * GW);                      // box [64 k][GW rows]
          } else {
#pragma unroll
            for (int nh = 0; nh < BNT / BN; ++nh)
#pragma unroll
              for (int bx = 0; bx < GW / 64; ++bx)
                load(sb + (nh * (GW / 64) + bx) * 8192, &tmB, w.n0 + nh * BN + rank * GW + 64 * bx, kb * BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA only) ===============================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc(A_MN ? 1 : 0, B_MN ? 1 : 0, BM * NCTA, BN);
      uint32_t stage = 0, phase = 0;
      uint32_t it = 0;
      TileLoop<MODE, CTA2> tl;
      tl.init(a, pid, npid);
      Work w;
      int res_m = -1;
      uint32_t aphase = 0;
      for (; tl.next(a, rank, w); ++it) {
        const uint32_t buf = it % NBUF, bphase = (it / NBUF) & 1;
        mbar_wait(bar_tempty + 8 * buf, bphase ^ 1);                       // epilogue(s) drained this accumulator
        if (AS && w.m_tile != res_m) {
          mbar_wait(bar_afull, aphase);
          aphase ^= 1;
          res_m = w.m_tile;
        }
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BN;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = AS ? (ares_base + kb * A_STAGE_BYTES) : (tiles_base + stage * STAGE_BYTES);
          const uint32_t sb = tiles_base + stage * STAGE_BYTES + A_IN_STAGE;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = A_MN ? desc_mnmajor(sa, k) : desc_kmajor(sa, k);
            const uint32_t acc = (kb > w.kb0 || k > 0) ? 1u : 0u;
#pragma unroll
            for (int nh = 0; nh < BNT / BN; ++nh) {                       // DX: two N=256 groups of the 512-wide tile
              const uint32_t sbh = sb + nh * (GW / 64) * 8192;
              const uint64_t db = B_MN ? desc_mnmajor(sbh, k) : desc_kmajor(sbh, k);
              if (CTA2) umma_bf16_2sm(tmem_d + nh * BN, da, db, idesc, acc);
              else umma_bf16(tmem_d + nh * BN, da, db, idesc, acc);
            }
          }
          // frees the smem stage (in both CTAs) when the MMAs retire
          if (CTA2) umma_commit_2sm(bar_empty + 8 * stage); else umma_commit(bar_empty + 8 * stage);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator ready for the epilogue(s)
        if (CTA2) umma_commit_2sm(bar_tfull + 8 * buf); else umma_commit(bar_tfull + 8 * buf);
        if (AS) { if (CTA2) umma_commit_2sm(bar_afree); else umma_commit(bar_afree); }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // =============================== epilogue ===============================
    // 8 warps: warp%4 selects the TMEM lane quarter (hardware rule), (warp-4)/4 the column half.
    const int q = warp & 3;
    const int half = (warp - EPI_WARP0) >> 2;
    const int r = q * 32 + lane;                    // row of the tile owned by this thread
    constexpr int NCHUNK = (BNT / 2) / 32;          // 32-column chunks per warp (4, DX: 8)
    const MhParams& p = a.p;
    const float ha = (V == V_CURR) ? a.state[4] : p.hard_a;
    const float hb = p.hard_b, lo = p.lo, hi = p.hi;
    uint32_t it = 0;
    TileLoop<MODE, CTA2> tl;
    tl.init(a, pid, npid);
    Work w;
    for (; tl.next(a, rank, w); ++it) {
      const uint32_t buf = it % NBUF, bphase = (it / NBUF) & 1;
      const int64_t row = (int64_t)w.m0 + r;
      RowCtx rc;
      rc.valid = true; rc.tcol = -1;
      if (MODE == MODE_FWD || MODE == MODE_BWD_G) {
        rc.valid = row < a.B;
        rc.scale = a.rowp[MH_RP_SCALE * a.ldp + row];
        rc.scale2 = rc.scale * MH_LOG2E;
        rc.thr = a.rowp[MH_RP_THR * a.ldp + row];
        rc.t = a.rowp[MH_RP_T * a.ldp + row];
        rc.zt2 = a.rowp[MH_RP_ZT * a.ldp + row] * MH_LOG2E;
        rc.dzt = a.rowp[MH_RP_DZT * a.ldp + row];
        const int32_t y = a.label_local[row];
        if (y >= w.n0 && y < w.n0 + BN) rc.tcol = y - w.n0;
        rc.lse2 = (MODE == MODE_BWD_G && rc.valid) ? a.lse2[row] : 0.f;
      }
      // DWF: this thread owns class `row`; dW_j = coef * (dw^_j - w^_j * rj)
      float dwf_rj = 0.f, dwf_coef = 0.f;
      bool dwf_ok = false;
      if (MODE == MODE_DWF) {
        dwf_ok = row < a.C;
        if (dwf_ok) {
          dwf_rj = a.rsum[row];
          dwf_coef = a.gscal[0] * a.inv_norm[row];
        }
      }
      const int nvalid = (int)min((int64_t)BN, a.C - (int64_t)w.n0);   // valid class columns in this tile (FWD/BWD_G)
      const int cbase = half * (BNT / 2);
      mbar_wait(bar_tfull + 8 * buf, bphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + cbase;

      FwdAcc acc{-INFINITY, 0.f, 0.f, 0};
      uint8_t* stg = stg_all + (warp - EPI_WARP0) * mode_stg_warp_bytes(MODE);   // this warp's staging rows
      uint32_t va[32], vb[32];
      tmem_ld32(taddr, va);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        uint32_t (&cur)[32] = (c & 1) ? vb : va;
        uint32_t (&nxt)[32] = (c & 1) ? va : vb;
        if (c + 1 < NCHUNK) tmem_ld32(taddr + (c + 1) * 32, nxt);   // prefetch while this chunk is processed
        const int col0 = cbase + c * 32;
        if (MODE == MODE_FWD) {
          fwd_chunk<V>(cur, col0, nvalid, rc, lo, hi, ha, hb, acc);
        } else if (MODE == MODE_BWD_G) {
          uint32_t pk[16];
          float qv[32];
          bwd_chunk<V>(cur, col0, nvalid, rc, lo, hi, ha, hb, pk, qv);
          // stage 64 B of this row: 16 B piece index XOR (row & 7) -> conflict-free writes and reads
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((((c & 1) * 4 + k) ^ (lane & 7)) * 16)) =
                make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
          if (a.rsum) {
            const float cs = warp_colsum32(qv, lane);                   // lane L: sum over this warp's 32 rows, column col0+L
            atomicAdd(a.rsum + w.n0 + col0 + lane, cs);
          }
          if (c & 1) {
            // two chunks staged = [32 rows][64 classes = 128 B] -> global as full lines (8 lanes per row, 4 rows per
            // instruction).  G is class-tiled [C_pad/128][B_pad][128]: this warp's 128 columns are one slab.
            __syncwarp();
            __nv_bfloat16* obase = a.G + (((int64_t)(w.n0 + cbase) / 128) * a.B_pad + w.m0 + q * 32) * 128 + (c - 1) * 32;
#pragma unroll
            for (int i2 = 0; i2 < 8; ++i2) {
              const int rr = 4 * i2 + (lane >> 3);
              const uint4 val = *reinterpret_cast<const uint4*>(stg + rr * 128 + (((lane & 7) ^ (rr & 7)) * 16));
              *reinterpret_cast<uint4*>(obase + (int64_t)rr * 128 + (lane & 7) * 8) = val;
            }
            __syncwarp();
          }
        } else if (MODE == MODE_DW || (MODE == MODE_DWF && a.layout == MH_LAYOUT_CD)) {
          float o[32];
          if (MODE == MODE_DWF) {
            const uint4* wsrc = reinterpret_cast<const uint4*>(a.w_hat + row * MH_D + w.n0 + col0);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              uint4 wq = dwf_ok ? __ldg(wsrc + k4) : make_uint4(0u, 0u, 0u, 0u);
              const uint32_t ww[4] = {wq.x, wq.y, wq.z, wq.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 wf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ww[e]));
                const int k = k4 * 8 + e * 2;
                o[k] = (__uint_as_float(cur[k]) - wf.x * dwf_rj) * dwf_coef;
                o[k + 1] = (__uint_as_float(cur[k + 1]) - wf.y * dwf_rj) * dwf_coef;
              }
            }
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) o[k] = __uint_as_float(cur[k]);
          }
          float4* dst = reinterpret_cast<float4*>(stg + lane * STG_ROW_BYTES + (c & 1) * 128);
#pragma unroll
          for (int k = 0; k < 8; ++k) dst[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
          if (c & 1) {
            // two chunks (64 fp32 columns = 256 B per row) staged: write them out as full 128 B lines,
            // 16 lanes per row, 2 rows per store instruction
            __syncwarp();
            const int64_t opitch = (MODE == MODE_DWF) ? a.ld : (int64_t)MH_D;
            const int64_t rbase = (int64_t)w.m0 + q * 32;
            float* obase = a.out + rbase * opitch + w.n0 + cbase + (c - 1) * 32;
#pragma unroll
            for (int i2 = 0; i2 < 16; ++i2) {
              const int rr = 2 * i2 + (lane >> 4);
              const uint4 val = *reinterpret_cast<const uint4*>(stg + rr * STG_ROW_BYTES + (lane & 15) * 16);
              if (MODE == MODE_DW || rbase + rr < a.C)
                *reinterpret_cast<uint4*>(obase + (int64_t)rr * opitch + (lane & 15) * 4) = val;
            }
            __syncwarp();
          }
        } else if (MODE == MODE_DWF) {
          // parameter layout [D, C]: for a fixed d the 32 lanes are 32 consecutive classes -> coalesced directly
          const uint4* wsrc = reinterpret_cast<const uint4*>(a.w_hat + row * MH_D + w.n0 + col0);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            uint4 wq = dwf_ok ? __ldg(wsrc + k4) : make_uint4(0u, 0u, 0u, 0u);
            const uint32_t ww[4] = {wq.x, wq.y, wq.z, wq.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 wf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ww[e]));
              const int k = k4 * 8 + e * 2;
              if (dwf_ok) {
                a.out[(int64_t)(w.n0 + col0 + k) * a.ld + row] = (__uint_as_float(cur[k]) - wf.x * dwf_rj) * dwf_coef;
                a.out[(int64_t)(w.n0 + col0 + k + 1) * a.ld + row] = (__uint_as_float(cur[k + 1]) - wf.y * dwf_rj) * dwf_coef;
              }
            }
          }
        } else {
          float* dst = a.out + (int64_t)w.split * a.out_split_stride + row * MH_D + w.n0 + col0;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            reinterpret_cast<uint4*>(dst)[k] = make_uint4(cur[4 * k], cur[4 * k + 1], cur[4 * k + 2], cur[4 * k + 3]);
        }
        if (c + 1 < NCHUNK) tmem_ld_wait();
      }
      // all TMEM reads of this accumulator are complete -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTA2) mbar_arrive_cluster(bar_tempty + 8 * buf, 0); else mbar_arrive(bar_tempty + 8 * buf);
      }
      if (MODE == MODE_FWD) {
        float* sp = a.stats_tiles + ((int64_t)w.n_tile * 2 + half) * MH_ST_PLANES * a.B_pad;
        sp[MH_ST_M * a.B_pad + row] = acc.m;
        sp[MH_ST_L * a.B_pad + row] = acc.l;
        sp[MH_ST_CNT * a.B_pad + row] = (float)acc.cnt;
        sp[MH_ST_EZ * a.B_pad + row] = (V == V_SPHERE) ? acc.ez : 0.f;
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();     // pair: the peer's smem / TMEM stay valid until both are done
  if (warp == 2) {
    tc_fence_after();
    if (CTA2) tmem_dealloc_2sm(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
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

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// cta_group::2 is the default; MH_TC_CTA2=0 selects the single-CTA kernels (kept for A/B measurements).
bool use_cta2() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MH_TC_CTA2");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <int MODE, int V, bool CTA2>
int launch_impl(const CUtensorMap& ta, const CUtensorMap& tb, const TcArgs& args, cudaStream_t st) {
  static bool attr_set = false;
  constexpr int smem = mode_smem_bytes(MODE, CTA2);
  if (!attr_set) {
    MH_CUDA_OK(cudaFuncSetAttribute(tc_kernel<MODE, V, CTA2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int units = CTA2 ? num_sms() / 2 : num_sms();
  // A-stationary kernels use a static schedule over exactly `units` pairs (pairs without work exit at once)
  const int n = mode_astat(MODE, CTA2) ? units : (int)std::min<int64_t>(args.total_tiles, units);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(CTA2 ? 2 * n : n);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTA2 ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MH_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_kernel<MODE, V, CTA2>, ta, tb, args));
  return MH_OK;
}

template <int MODE, int V>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const TcArgs& args, bool cta2, cudaStream_t st) {
  return cta2 ? launch_impl<MODE, V, true>(ta, tb, args, st) : launch_impl<MODE, V, false>(ta, tb, args, st);
}

int variant_of(const MhParams& p) {
  if (p.hard_kind == 1) return V_MV;
  if (p.hard_kind == 2) return V_CURR;
  if (p.scale_is_norm) return V_SPHERE;
  if (p.family == MH_ARCFACE) return V_PLAIN;
  return V_CLAMP;
}

template <int MODE>
int launch_variant(const CUtensorMap& ta, const CUtensorMap& tb, const TcArgs& args, bool cta2, cudaStream_t st) {
  switch (variant_of(args.p)) {
    case V_PLAIN: return launch<MODE, V_PLAIN>(ta, tb, args, cta2, st);
    case V_CLAMP: return launch<MODE, V_CLAMP>(ta, tb, args, cta2, st);
    case V_SPHERE: return launch<MODE, V_SPHERE>(ta, tb, args, cta2, st);
    case V_MV: return launch<MODE, V_MV>(ta, tb, args, cta2, st);
    default: return launch<MODE, V_CURR>(ta, tb, args, cta2, st);
  }
}

}  // namespace

// two statistics records per 256-wide class tile (one per 128-column epilogue half)
extern "C" int64_t mh_fwd_num_tiles(int64_t C_pad) { return 2 * ((C_pad + BN - 1) / BN); }

static int check_common(int64_t B, int64_t B_pad, int64_t C, int64_t C_pad) {
  MH_CHECK_ARG(B > 0 && B_pad >= B && B_pad % BM == 0, "B_pad must be a multiple of 128");
  MH_CHECK_ARG(C > 0 && C_pad >= C && C_pad % BN == 0, "C_pad must be a multiple of 256 for the tensor-core path");
  MH_CHECK_ARG(C_pad < (1ll << 31) && B_pad < (1ll << 31) && (C_pad / 128) * B_pad < (1ll << 31), "dimension too large");
  return MH_OK;
}

// A-stationary schedule parameters for m_tiles <= units row tiles (see StatIter).
static void make_sched(TcArgs& a, int units) {
  a.sG = units / a.m_tiles;
  a.sE = units - a.sG * a.m_tiles;
  const int64_t wfix = (int64_t)a.sG * a.m_tiles;
  a.n_fixed = a.sE == 0 ? a.n_tiles : (int)(((int64_t)a.n_tiles * wfix + (wfix + a.sE) / 2) / (wfix + a.sE));
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

// FWD / BWD_G launch.  cta2: A-stationary schedule; row tiles beyond `units` pairs go in further launches.
template <int MODE>
static int launch_s_tiles(const mh_config* cfg_host, const void* x_hat_bf16, int64_t B, int64_t B_pad, const void* w_hat_bf16,
                          int64_t C, int64_t C_pad, const float* rowp, int64_t ldp, const int32_t* label_local,
                          const float* state, const float* lse2, float* stats_tiles, void* G_bf16, float* r_colsum,
                          cudaStream_t st) {
  const bool cta2 = use_cta2() && (B_pad % (2 * BM) == 0);
  const int bmt = cta2 ? 2 * BM : BM;
  const int units = cta2 ? num_sms() / 2 : num_sms();
  CUtensorMap tb;
  if (int e = make_tmap(&tb, w_hat_bf16, C_pad, MH_D, cta2 ? BN / 2 : BN)) return e;
  const int64_t rows_per_launch = cta2 ? (int64_t)units * bmt : B_pad;
  for (int64_t r0 = 0; r0 < B_pad; r0 += rows_per_launch) {
    const int64_t rows = std::min(rows_per_launch, B_pad - r0);
    if (r0 >= B) break;                                               // only padding rows left
    CUtensorMap ta;
    if (int e = make_tmap(&ta, (const __nv_bfloat16*)x_hat_bf16 + r0 * MH_D, rows, MH_D, BM)) return e;
    TcArgs a{};
    a.m_tiles = (int)(rows / bmt); a.n_tiles = (int)(C_pad / BN); a.n_split = 1;
    a.k_blocks_total = MH_D / BK; a.k_blocks_per_split = a.k_blocks_total;
    a.total_tiles = (int64_t)a.m_tiles * a.n_tiles;
    if (cta2) make_sched(a, units);
    a.p = mh_make_params(cfg_host);
    a.B = B - r0; a.C = C; a.B_pad = B_pad; a.C_pad = C_pad;
    a.rowp = rowp + r0; a.ldp = ldp; a.label_local = label_local + r0; a.state = state;
    a.lse2 = lse2 ? lse2 + r0 : nullptr;
    a.stats_tiles = stats_tiles ? stats_tiles + r0 : nullptr;
    a.G = G_bf16 ? (__nv_bfloat16*)G_bf16 + r0 * 128 : nullptr;
    a.rsum = r_colsum;
    if (int e = launch_variant<MODE>(ta, tb, a, cta2, st)) return e;
  }
  return MH_OK;
}

extern "C" int mh_tc_forward(const mh_config* cfg_host, const void* x_hat_bf16, int64_t B, int64_t B_pad,
                             const void* w_hat_bf16, int64_t C, int64_t C_pad, const float* rowp, int64_t ldp,
                             const int32_t* label_local, const float* state, float* stats_tiles, void* stream) {
  MH_CHECK_ARG(cfg_host && x_hat_bf16 && w_hat_bf16 && rowp && label_local && state && stats_tiles, "null pointer");
  if (int e = check_common(B, B_pad, C, C_pad)) return e;
  MH_CHECK_ARG(ldp >= B_pad, "rowp pitch must cover B_pad");
  return launch_s_tiles<MODE_FWD>(cfg_host, x_hat_bf16, B, B_pad, w_hat_bf16, C, C_pad, rowp, ldp, label_local, state,
                                  nullptr, stats_tiles, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int mh_tc_backward_g(const mh_config* cfg_host, const void* x_hat_bf16, int64_t B, int64_t B_pad,
                                const void* w_hat_bf16, int64_t C, int64_t C_pad, const float* rowp, int64_t ldp,
                                const int32_t* label_local, const float* state, const float* lse2, void* G_bf16,
                                float* r_colsum, void* stream) {
  MH_CHECK_ARG(cfg_host && x_hat_bf16 && w_hat_bf16 && rowp && label_local && state && lse2 && G_bf16, "null pointer");
  if (int e = check_common(B, B_pad, C, C_pad)) return e;
  MH_CHECK_ARG(ldp >= B_pad, "rowp pitch must cover B_pad");
  if (r_colsum) MH_CUDA_OK(cudaMemsetAsync(r_colsum, 0, sizeof(float) * C_pad, (cudaStream_t)stream));
  return launch_s_tiles<MODE_BWD_G>(cfg_host, x_hat_bf16, B, B_pad, w_hat_bf16, C, C_pad, rowp, ldp, label_local, state,
                                    lse2, nullptr, G_bf16, r_colsum, (cudaStream_t)stream);
}

extern "C" int mh_tc_backward_dx(const void* G_bf16, int64_t B_pad, int64_t C_pad, const void* w_hat_bf16, float* out,
                                 int* n_split_host, void* stream) {
  MH_CHECK_ARG(B_pad > 0 && B_pad % BM == 0 && C_pad > 0 && C_pad % BN == 0, "bad padded shape");
  const bool cta2 = use_cta2() && (B_pad % (2 * BM) == 0);
  const int m_tiles = (int)(B_pad / (cta2 ? 2 * BM : BM));
  const int kb_total = (int)(C_pad / BK);
  int n_split = std::max(1, (cta2 ? num_sms() / 2 : num_sms()) / m_tiles);
  n_split = std::min(n_split, kb_total);
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
  return launch<MODE_DX, V_NONE>(ta, tb, a, cta2, (cudaStream_t)stream);
}

extern "C" int mh_tc_backward_dw(const void* G_bf16, int64_t B_pad, int64_t C_pad, const void* x_hat_bf16,
                                 float* dw_hat, void* stream) {
  MH_CHECK_ARG(G_bf16 && x_hat_bf16 && dw_hat, "null pointer");
  MH_CHECK_ARG(B_pad > 0 && B_pad % BM == 0 && C_pad > 0 && C_pad % BN == 0, "bad padded shape");
  CUtensorMap ta, tb;
  if (int e = make_tmap(&ta, G_bf16, (C_pad / 128) * B_pad, 128, 64)) return e;    // A = G^T (class-tiled G), MN-major boxes
  if (int e = make_tmap(&tb, x_hat_bf16, B_pad, MH_D, 64)) return e;       // B = x^,  MN-major boxes [64 rows][64 d]
  const bool cta2 = use_cta2();                                     // C_pad is always a multiple of 256
  TcArgs a{};
  a.m_tiles = (int)(C_pad / (cta2 ? 2 * BM : BM)); a.n_tiles = 2; a.n_split = 1;
  a.k_blocks_total = (int)(B_pad / BK); a.k_blocks_per_split = a.k_blocks_total;
  a.total_tiles = (int64_t)a.m_tiles * 2;
  a.B_pad = B_pad; a.C_pad = C_pad; a.B = B_pad; a.C = C_pad;
  a.out = dw_hat; a.out_split_stride = 0;
  return launch<MODE_DW, V_NONE>(ta, tb, a, cta2, (cudaStream_t)stream);
}

extern "C" int mh_tc_backward_dw_fused(const void* G_bf16, int64_t B_pad, int64_t C, int64_t C_pad, const void* x_hat_bf16,
                                       const void* w_hat_bf16, const float* inv_norm, const float* r_colsum,
                                       const float* gscal, int layout, float* dW, int64_t ld, void* stream) {
  MH_CHECK_ARG(G_bf16 && x_hat_bf16 && w_hat_bf16 && inv_norm && r_colsum && gscal && dW, "null pointer");
  MH_CHECK_ARG(B_pad > 0 && B_pad % BM == 0 && C_pad > 0 && C_pad % BN == 0 && C > 0 && C <= C_pad, "bad padded shape");
  MH_CHECK_ARG(layout == MH_LAYOUT_CD || layout == MH_LAYOUT_DC, "unknown layout");
  MH_CHECK_ARG(layout != MH_LAYOUT_CD || (ld % 4 == 0 && ((uintptr_t)dW & 15) == 0), "CD dW must be 16-byte aligned");
  CUtensorMap ta, tb;
  if (int e = make_tmap(&ta, G_bf16, (C_pad / 128) * B_pad, 128, 64)) return e;    // class-tiled G
  if (int e = make_tmap(&tb, x_hat_bf16, B_pad, MH_D, 64)) return e;
  const bool cta2 = use_cta2();
  TcArgs a{};
  a.m_tiles = (int)(C_pad / (cta2 ? 2 * BM : BM)); a.n_tiles = 2; a.n_split = 1;
  a.k_blocks_total = (int)(B_pad / BK); a.k_blocks_per_split = a.k_blocks_total;
  a.total_tiles = (int64_t)a.m_tiles * 2;
  a.B_pad = B_pad; a.C_pad = C_pad; a.B = B_pad; a.C = C;
  a.out = dW; a.out_split_stride = 0;
  a.rsum = const_cast<float*>(r_colsum); a.w_hat = (const __nv_bfloat16*)w_hat_bf16; a.inv_norm = inv_norm;
  a.gscal = gscal; a.layout = layout; a.ld = ld;
  return launch<MODE_DWF, V_NONE>(ta, tb, a, cta2, (cudaStream_t)stream);
}
