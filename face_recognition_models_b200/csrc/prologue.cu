// HBM-bound prologues: L2-normalise class centres and embeddings, gather the target cosine,
// and compute the per-row margin terms.  Replaces F.normalize / torch.norm / the target gather
// of the reference heads (criterion.py:65,95,173-174,263-264,400-404,417,538-542,552,860-864,...).
#include "common.cuh"
#include <cstdlib>

// ------------------------------------------------------------------------------------------------
// Optional fused optimizer: the reference trains the class centres with SGD(momentum 0.9, weight_decay 5e-4)
// (model_utils.py:557).  The SGD variants of the kernels below apply that update to W while they stream it, so the
// optimizer pass also leaves the next step's w^ / inv_norm behind and the next forward skips prologue_w.
// Per element (torch.optim.SGD, dampening 0, no nesterov):  g' = g / scale + wd * w;  buf = momentum * buf + g';
// w -= lr * buf.  A non-zero *found_inf (GradScaler) turns the update off; w^ is then rebuilt from the unchanged W.
// ------------------------------------------------------------------------------------------------
struct SgdArgs {
  const float* grad;
  float* mom;
  float lr, momentum, wd;
  const float* grad_scale;   // device scalar or null
  const float* found_inf;    // device scalar or null
};

struct SgdCtx {
  bool live;
  float inv_scale;
};

__device__ __forceinline__ SgdCtx sgd_ctx(const SgdArgs& s) {
  SgdCtx c;
  c.live = s.found_inf == nullptr || __ldg(s.found_inf) == 0.f;
  c.inv_scale = s.grad_scale ? 1.f / __ldg(s.grad_scale) : 1.f;
  return c;
}

__device__ __forceinline__ float sgd_elem(float w, float g, float& m, const SgdArgs& s, const SgdCtx& c) {
  const float gp = fmaf(s.wd, w, g * c.inv_scale);
  m = __fadd_rn(__fmul_rn(m, s.momentum), gp);
  return fmaf(-s.lr, m, w);
}

__device__ __forceinline__ float4 sgd_elem4(float4 w, float4 g, float4& m, const SgdArgs& s, const SgdCtx& c) {
  float4 o;
  o.x = sgd_elem(w.x, g.x, m.x, s, c);
  o.y = sgd_elem(w.y, g.y, m.y, s, c);
  o.z = sgd_elem(w.z, g.z, m.z, s, c);
  o.w = sgd_elem(w.w, g.w, m.w, s, c);
  return o;
}

// ------------------------------------------------------------------------------------------------
// prologue_w, layout CD: W [C, 512] row-major. One warp per class; 16 B vector loads, 8 B bf16 stores.
// Algorithmic bytes per class: 2048 read + 1024 bf16 write (+2048 optional fp32 copy) + 4.
// SGD variant: + 4096 read (grad, momentum) + 4096 write (W, momentum).
// ------------------------------------------------------------------------------------------------
template <bool SGD>
__global__ void __launch_bounds__(256) prologue_w_cd_kernel(float* __restrict__ W, int64_t C, int64_t ld,
                                                            __nv_bfloat16* __restrict__ what, int64_t C_pad,
                                                            float* __restrict__ what32, float* __restrict__ inv_norm,
                                                            SgdArgs sg) {
  mh_pdl_sync();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= C_pad) return;
  uint2* dst = reinterpret_cast<uint2*>(what + row * MH_D);
  if (row >= C) {
#pragma unroll
    for (int k = 0; k < 4; ++k) dst[lane + 32 * k] = make_uint2(0u, 0u);
    return;
  }
  float4* src = reinterpret_cast<float4*>(W + row * ld);
  float4 v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = SGD ? src[lane + 32 * k] : __ldg(src + lane + 32 * k);
  if (SGD) {
    const SgdCtx cx = sgd_ctx(sg);
    if (cx.live) {
      const float4* gsrc = reinterpret_cast<const float4*>(sg.grad + row * ld);
      float4* msrc = reinterpret_cast<float4*>(sg.mom + row * ld);
      float4 g[4], m[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        g[k] = __ldg(gsrc + lane + 32 * k);
        m[k] = msrc[lane + 32 * k];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[k] = sgd_elem4(v[k], g[k], m[k], sg, cx);
        src[lane + 32 * k] = v[k];
        msrc[lane + 32 * k] = m[k];
      }
    }
  }
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) ss += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  if (lane == 0) inv_norm[row] = inv;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float4 o = make_float4(v[k].x * inv, v[k].y * inv, v[k].z * inv, v[k].w * inv);
    __nv_bfloat162 p0 = __floats2bfloat162_rn(o.x, o.y), p1 = __floats2bfloat162_rn(o.z, o.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&p0);
    pk.y = *reinterpret_cast<uint32_t*>(&p1);
    dst[lane + 32 * k] = pk;
    if (what32) reinterpret_cast<float4*>(what32 + row * MH_D)[lane + 32 * k] = o;
  }
}

// ------------------------------------------------------------------------------------------------
// prologue_w, layout DC: W [512, C] (class index contiguous). A block stages a [512 x 32-class] slab
// in shared memory with coalesced 128 B row reads, reduces column norms, and writes the transposed,
// normalised bf16 rows (coalesced along d). One read of W, one write of w_hat.
// ------------------------------------------------------------------------------------------------
#define PW_TC 32
template <bool SGD>
__global__ void __launch_bounds__(256) prologue_w_dc_kernel(float* __restrict__ W, int64_t C, int64_t ld,
                                                            __nv_bfloat16* __restrict__ what, int64_t C_pad,
                                                            float* __restrict__ what32, float* __restrict__ inv_norm,
                                                            SgdArgs sg) {
  mh_pdl_sync();
  extern __shared__ float slab[];            // [512][33]
  __shared__ float part[8][PW_TC];
  __shared__ float invs[PW_TC];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c0 = (int64_t)blockIdx.x * PW_TC;
  const int64_t c = c0 + tx;
  float ss = 0.f;
  // 16 independent 128-byte row loads in flight per warp (the loop was latency-bound at the compiler's unroll of 4:
  // 2.6 TB/s; HBM needs ~40 KB in flight per SM); the SGD variant has 3 arrays to load, 8 rows of each per batch
  constexpr int U = SGD ? 8 : 16;
  SgdCtx cx;
  if (SGD) cx = sgd_ctx(sg);
#pragma unroll 1
  for (int d0 = ty; d0 < MH_D; d0 += 8 * U) {
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float* src = W + (int64_t)(d0 + 8 * u) * ld + c;
      v[u] = (c < C) ? (SGD ? *src : __ldg(src)) : 0.f;
    }
    if (SGD) {
      if (cx.live && c < C) {
        float g[U], m[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t off = (int64_t)(d0 + 8 * u) * ld + c;
          g[u] = __ldg(sg.grad + off);
          m[u] = sg.mom[off];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t off = (int64_t)(d0 + 8 * u) * ld + c;
          v[u] = sgd_elem(v[u], g[u], m[u], sg, cx);
          W[off] = v[u];
          sg.mom[off] = m[u];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) ss += v[u] * v[u];
    const int skew = (ss < 0.f) ? 1 : 0;     // always 0; ties the stores to the whole batch of loads (see the dc4 kernel)
#pragma unroll
    for (int u = 0; u < U; ++u) slab[(d0 + 8 * u) * 33 + tx + skew] = v[u];
  }
  part[ty][tx] = ss;
  __syncthreads();
  if (ty == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][tx];
    float inv = 1.f / fmaxf(sqrtf(t), 1e-12f);
    invs[tx] = inv;
    if (c < C) inv_norm[c] = inv;
  }
  __syncthreads();
  // each warp writes 4 class rows; lanes run along d (4 consecutive d per lane per step)
  for (int r = ty; r < PW_TC; r += 8) {
    const int64_t row = c0 + r;
    if (row >= C_pad) continue;
    const float inv = (row < C) ? invs[r] : 0.f;
    // two consecutive d per lane: the column read of the [d][33] slab is 2-way bank-conflicted (4-way with four),
    // and a warp still stores full 128-byte lines of bf16
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int d = 2 * tx + 64 * k;
      const float2 o = make_float2(slab[d * 33 + r] * inv, slab[(d + 1) * 33 + r] * inv);
      *reinterpret_cast<__nv_bfloat162*>(what + row * MH_D + d) = __floats2bfloat162_rn(o.x, o.y);
      if (what32 && row < C) *reinterpret_cast<float2*>(what32 + row * MH_D + d) = o;
    }
  }
}

// Same slab kernel for 16-byte-aligned rows (ld % 4 == 0, e.g. C = 2,000,000): one LDG.128 covers 4 d-rows x 32 classes
// (8 lanes x 16 B per row), 8 of them in flight per warp = 4 KB, 96 KB per SM at 3 resident blocks.
template <bool SGD>
__global__ void __launch_bounds__(256) prologue_w_dc4_kernel(float* __restrict__ W, int64_t C, int64_t ld,
                                                             __nv_bfloat16* __restrict__ what, int64_t C_pad,
                                                             float* __restrict__ what32, float* __restrict__ inv_norm,
                                                             SgdArgs sg) {
  mh_pdl_sync();
  // [512 d][8 chunks of 4 classes], 128 B per d-row; chunk q of row d sits at position q ^ ((d >> 1) & 7), which makes
  // both the 16-byte stores of the load phase (8 lanes = 8 chunks of one row) and the 16-byte column reads of the write
  // phase (8 lanes = one chunk of 8 row pairs) conflict-free.  (4-byte accesses into a [512][33] slab kept the
  // shared-memory instruction queue full: mio_throttle was the top stall, 1.38 ms at C = 2M.)
  extern __shared__ float4 slab4[];
  __shared__ float part[8][PW_TC];
  __shared__ __align__(16) float invs[PW_TC];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r4 = tx >> 3, q = tx & 7;        // row within the group of 4, float4 index within the 32-class row segment
  const int64_t c0 = (int64_t)blockIdx.x * PW_TC;
  const int64_t cq = c0 + 4 * q;
  float ss4[4] = {0.f, 0.f, 0.f, 0.f};
  constexpr int U = SGD ? 8 : 16;            // plain: all 16 row groups of the warp in flight at once (8 KB per warp)
  SgdCtx cx;
  if (SGD) cx = sgd_ctx(sg);
#pragma unroll 1
  for (int it0 = 0; it0 < 16; it0 += U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int d = 4 * (ty + 8 * (it0 + u)) + r4;
      // C % 4 == 0, so a float4 is either entirely inside [0, C) or entirely outside
      const float4* src = reinterpret_cast<const float4*>(W + (int64_t)d * ld + cq);
      v[u] = (cq < C) ? (SGD ? *src : __ldg(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (SGD) {
      if (cx.live && cq < C) {
        float4 g[U], m[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t off = (int64_t)(4 * (ty + 8 * (it0 + u)) + r4) * ld + cq;
          g[u] = __ldg(reinterpret_cast<const float4*>(sg.grad + off));
          m[u] = *reinterpret_cast<const float4*>(sg.mom + off);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t off = (int64_t)(4 * (ty + 8 * (it0 + u)) + r4) * ld + cq;
          v[u] = sgd_elem4(v[u], g[u], m[u], sg, cx);
          *reinterpret_cast<float4*>(W + off) = v[u];
          *reinterpret_cast<float4*>(sg.mom + off) = m[u];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      ss4[0] += v[u].x * v[u].x; ss4[1] += v[u].y * v[u].y; ss4[2] += v[u].z * v[u].z; ss4[3] += v[u].w * v[u].w;
    }
    // ptxas otherwise sinks each load next to its shared-memory store (3 loads in flight instead of the whole batch):
    // the store address below depends on the finished sums, i.e. on every load of the batch.  `skew` is always 0.
    const int skew = (ss4[0] + ss4[1] + ss4[2] + ss4[3] < 0.f) ? 1 : 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int d = 4 * (ty + 8 * (it0 + u)) + r4;
      slab4[d * 8 + (q ^ ((d >> 1) & 7)) + skew] = v[u];
    }
  }
  // sum over the 4 row groups of the warp (lane bits 3, 4), then over the 8 warps
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    ss4[e] += __shfl_xor_sync(0xffffffffu, ss4[e], 8);
    ss4[e] += __shfl_xor_sync(0xffffffffu, ss4[e], 16);
  }
  if (r4 == 0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) part[ty][4 * q + e] = ss4[e];
  }
  __syncthreads();
  const int64_t c = c0 + tx;
  if (ty == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][tx];
    float inv = 1.f / fmaxf(sqrtf(t), 1e-12f);
    invs[tx] = inv;
    if (c < C) inv_norm[c] = inv;
  }
  __syncthreads();
  // write phase: a lane takes 4 classes x 2 consecutive d (two 16-byte reads), a warp stores 128-byte lines of bf16
  for (int it = ty; it < 64; it += 8) {
    const int j = it & 7, k = it >> 3;         // class chunk, 64-wide d block
    const int d = 64 * k + 2 * tx;
    const int pos = j ^ ((d >> 1) & 7);
    const float4 a = slab4[d * 8 + pos], b = slab4[(d + 1) * 8 + pos];
    const float4 iv = *reinterpret_cast<const float4*>(invs + 4 * j);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w}, ivv[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int64_t row = c0 + 4 * j + e;
      if (row >= C_pad) continue;
      const float inv = (row < C) ? ivv[e] : 0.f;
      const float2 o = make_float2(av[e] * inv, bv[e] * inv);
      *reinterpret_cast<__nv_bfloat162*>(what + row * MH_D + d) = __floats2bfloat162_rn(o.x, o.y);
      if (what32 && row < C) *reinterpret_cast<float2*>(what32 + row * MH_D + d) = o;
    }
  }
}

// Plain prologue (no optimizer), 16-byte-aligned DC rows: persistent blocks, one per SM, stream 32-class tiles through a
// 3-deep ring of swizzled slabs with cp.async, so two tiles (128 KB per SM) are always in flight while the block
// reduces and writes the third.  (The one-tile-per-block kernel above alternates load / barrier / write phases with
// 24 resident warps and stalls on long_scoreboard + barrier: 4.5 TB/s at C = 2M; this one 5.1-5.6 TB/s.)  Thread -> chunk mapping, summation order and
// therefore every output bit are those of prologue_w_dc4_kernel.
#define PW_NST 3
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, int src_bytes) {
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gsrc), "r"(src_bytes) : "memory");
}

__global__ void __launch_bounds__(256, 1) prologue_w_dc4p_kernel(const float* __restrict__ W, int64_t C, int64_t ld,
                                                                 __nv_bfloat16* __restrict__ what, int64_t C_pad,
                                                                 float* __restrict__ what32, float* __restrict__ inv_norm,
                                                                 int64_t n_tiles) {
  mh_pdl_sync();
  extern __shared__ float4 slab4[];            // PW_NST x [512 d][8 chunks], chunk q of row d at q ^ ((d >> 1) & 7)
  __shared__ float part[8][PW_TC];
  __shared__ __align__(16) float invs[PW_TC];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r4 = tx >> 3, q = tx & 7;
  auto issue = [&](int64_t tile, int buf) {
    const int64_t cq = tile * PW_TC + 4 * q;
    const bool in = cq < C;                    // C % 4 == 0: a chunk is entirely inside or outside
    const float* src = W + (in ? cq : 0);
    float4* dst = slab4 + buf * (MH_D * 8);
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int d = 4 * (ty + 8 * u) + r4;
      cp_async16_zfill(dst + d * 8 + (q ^ ((d >> 1) & 7)), src + (int64_t)d * ld, in ? 16 : 0);
    }
  };
  int64_t t = blockIdx.x;
#pragma unroll
  for (int s = 0; s < PW_NST - 1; ++s) {
    if (t + (int64_t)s * gridDim.x < n_tiles) issue(t + (int64_t)s * gridDim.x, s);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  int buf = 0;
  for (; t < n_tiles; t += gridDim.x) {
    const int64_t tn = t + (int64_t)(PW_NST - 1) * gridDim.x;
    if (tn < n_tiles) issue(tn, (buf + PW_NST - 1) % PW_NST);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(PW_NST - 1) : "memory");
    const float4* cur = slab4 + buf * (MH_D * 8);
    const int64_t c0 = t * PW_TC;
    float ss4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 16; ++u) {             // the thread's own 16 chunks, in the order of the register kernel
      const int d = 4 * (ty + 8 * u) + r4;
      const float4 v = cur[d * 8 + (q ^ ((d >> 1) & 7))];
      ss4[0] += v.x * v.x; ss4[1] += v.y * v.y; ss4[2] += v.z * v.z; ss4[3] += v.w * v.w;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      ss4[e] += __shfl_xor_sync(0xffffffffu, ss4[e], 8);
      ss4[e] += __shfl_xor_sync(0xffffffffu, ss4[e], 16);
    }
    if (r4 == 0) {
#pragma unroll
      for (int e = 0; e < 4; ++e) part[ty][4 * q + e] = ss4[e];
    }
    __syncthreads();                           // every thread has waited for its own copies: the slab is complete
    if (ty == 0) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += part[k][tx];
      const float inv = 1.f / fmaxf(sqrtf(acc), 1e-12f);
      invs[tx] = inv;
      if (c0 + tx < C) inv_norm[c0 + tx] = inv;
    }
    __syncthreads();
#pragma unroll 2
    for (int it = ty; it < 64; it += 8) {
      const int j = it & 7, k = it >> 3;       // class chunk (4 classes, all inside or all outside [0, C)), 64-wide d block
      const int d = 64 * k + 2 * tx;
      const int pos = j ^ ((d >> 1) & 7);
      const float4 a = cur[d * 8 + pos], b = cur[(d + 1) * 8 + pos];
      const int64_t row0 = c0 + 4 * j;
      if (row0 >= C_pad) continue;
      const bool in = row0 < C;
      float4 iv = *reinterpret_cast<const float4*>(invs + 4 * j);
      if (!in) iv = make_float4(0.f, 0.f, 0.f, 0.f);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w}, ivv[4] = {iv.x, iv.y, iv.z, iv.w};
      __nv_bfloat16* dst = what + row0 * MH_D + d;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 o = make_float2(av[e] * ivv[e], bv[e] * ivv[e]);
        *reinterpret_cast<__nv_bfloat162*>(dst + e * MH_D) = __floats2bfloat162_rn(o.x, o.y);
        if (what32 && in) *reinterpret_cast<float2*>(what32 + (row0 + e) * MH_D + d) = o;
      }
    }
    __syncthreads();                           // the slab, part[] and invs[] are reused from here on
    buf = (buf + 1) % PW_NST;
  }
}

// zero the padding rows [C, C_pad) that no DC block covers
__global__ void zero_pad_rows_kernel(__nv_bfloat16* what, int64_t row0, int64_t row1) {
  mh_pdl_sync();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t n = (row1 - row0) * MH_D;
  if (i < n) what[row0 * MH_D + i] = __float2bfloat16(0.f);
}

template <bool SGD>
static int launch_prologue_w(float* W, int layout, int64_t C, int64_t ld, void* w_hat_bf16, int64_t C_pad, float* w_hat32,
                             float* inv_norm, const SgdArgs& sg, cudaStream_t st) {
  if (layout == MH_LAYOUT_CD) {
    MH_CHECK_ARG(ld >= MH_D && ld % 4 == 0, "CD layout needs ld >= 512 and ld % 4 == 0");
    dim3 grid((unsigned)((C_pad + 7) / 8));
    mh_launch(prologue_w_cd_kernel<SGD>, grid, 256, 0, st, W, C, ld, (__nv_bfloat16*)w_hat_bf16, C_pad, w_hat32, inv_norm, sg);
  } else if (layout == MH_LAYOUT_DC) {
    MH_CHECK_ARG(ld >= C, "DC layout needs ld >= C");
    static MhDeviceOnce attr_once;
    const int smem = MH_D * 33 * sizeof(float);
    const int smem_p = PW_NST * MH_D * 8 * (int)sizeof(float4);
    MH_CUDA_OK(mh_once_per_device(attr_once, [&] {
      cudaError_t e = cudaFuncSetAttribute(prologue_w_dc_kernel<SGD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(prologue_w_dc4_kernel<SGD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(prologue_w_dc4p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_p);
      return e;
    }));
    dim3 grid((unsigned)((C + PW_TC - 1) / PW_TC));
    bool vec4 = ld % 4 == 0 && C % 4 == 0 && ((uintptr_t)W & 15) == 0;
    if (SGD) vec4 = vec4 && ((uintptr_t)sg.grad & 15) == 0 && ((uintptr_t)sg.mom & 15) == 0;
    if (vec4 && !SGD) {
      const int n_sm = mh_num_sms();
      const int64_t n_tiles = grid.x;
      const unsigned nb = (unsigned)(n_tiles < n_sm ? n_tiles : n_sm);
      mh_launch(prologue_w_dc4p_kernel, nb, 256, smem_p, st, W, C, ld, (__nv_bfloat16*)w_hat_bf16, C_pad, w_hat32, inv_norm, n_tiles);
    } else if (vec4)
      mh_launch(prologue_w_dc4_kernel<SGD>, grid, 256, smem, st, W, C, ld, (__nv_bfloat16*)w_hat_bf16, C_pad, w_hat32, inv_norm, sg);
    else
      mh_launch(prologue_w_dc_kernel<SGD>, grid, 256, smem, st, W, C, ld, (__nv_bfloat16*)w_hat_bf16, C_pad, w_hat32, inv_norm, sg);
    const int64_t covered = (int64_t)grid.x * PW_TC;
    if (covered < C_pad) {
      int64_t n = (C_pad - covered) * MH_D;
      mh_launch(zero_pad_rows_kernel, (unsigned)((n + 255) / 256), 256, 0, st, (__nv_bfloat16*)w_hat_bf16, covered, C_pad);
    }
  } else {
    MH_CHECK_ARG(false, "unknown layout");
  }
  MH_LAUNCH_OK();
  return MH_OK;
}

extern "C" int mh_prologue_w(const float* W, int layout, int64_t C, int64_t ld, void* w_hat_bf16, int64_t C_pad,
                             float* w_hat32, float* inv_norm, void* stream) {
  MH_CHECK_ARG(W && w_hat_bf16 && inv_norm, "null pointer");
  MH_CHECK_ARG(C > 0 && C_pad >= C && C_pad % MH_TILE == 0, "C_pad must be a multiple of 128 and >= C");
  MH_CHECK_ARG(((uintptr_t)W & 15) == 0 && ((uintptr_t)w_hat_bf16 & 15) == 0, "pointers must be 16-byte aligned");
  SgdArgs none{};
  return launch_prologue_w<false>(const_cast<float*>(W), layout, C, ld, w_hat_bf16, C_pad, w_hat32, inv_norm, none,
                                  (cudaStream_t)stream);
}

extern "C" int mh_sgd_step_w(float* W, int layout, int64_t C, int64_t ld, const float* grad, float* momentum_buf, float lr,
                             float momentum, float weight_decay, const float* grad_scale, const float* found_inf,
                             void* w_hat_bf16, int64_t C_pad, float* inv_norm, void* stream) {
  MH_CHECK_ARG(W && grad && momentum_buf && w_hat_bf16 && inv_norm, "null pointer");
  MH_CHECK_ARG(C > 0 && C_pad >= C && C_pad % MH_TILE == 0, "C_pad must be a multiple of 128 and >= C");
  MH_CHECK_ARG(((uintptr_t)W & 15) == 0 && ((uintptr_t)w_hat_bf16 & 15) == 0, "pointers must be 16-byte aligned");
  if (layout == MH_LAYOUT_CD)
    MH_CHECK_ARG(((uintptr_t)grad & 15) == 0 && ((uintptr_t)momentum_buf & 15) == 0, "pointers must be 16-byte aligned");
  MH_CHECK_ARG(lr >= 0.f && momentum >= 0.f && weight_decay >= 0.f, "lr, momentum and weight_decay must be >= 0");
  SgdArgs sg{grad, momentum_buf, lr, momentum, weight_decay, grad_scale, found_inf};
  return launch_prologue_w<true>(W, layout, C, ld, w_hat_bf16, C_pad, nullptr, inv_norm, sg, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// vpl_mix: v_j = (1 - a_j) w^_j + a_j m^_j, a_j = lamda * 1[life_j > 0]  (VPLArcFace, criterion.py:716-724).
// One warp per class: reads w^_j (bf16, 1 KB) and mem_j (fp32, 2 KB), writes v_j (bf16, 1 KB).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vpl_mix_kernel(const __nv_bfloat16* __restrict__ what, const float* __restrict__ mem,
                                                      const float* __restrict__ life, float lamda, int64_t C, int64_t C_pad,
                                                      __nv_bfloat16* __restrict__ v, float* __restrict__ alpha) {
  mh_pdl_sync();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= C_pad) return;
  uint2* dst = reinterpret_cast<uint2*>(v + row * MH_D);
  if (row >= C) {
#pragma unroll
    for (int k = 0; k < 4; ++k) dst[lane + 32 * k] = make_uint2(0u, 0u);
    return;
  }
  const float al = (life[row] > 0.f) ? lamda : 0.f;        // fl32(mask * lamda), as the reference's float32 mask
  const float be = 1.f - al;
  if (lane == 0) alpha[row] = al;
  float4 m4[4];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    m4[k] = (al != 0.f) ? __ldg(reinterpret_cast<const float4*>(mem + row * MH_D) + lane + 32 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
    ss += m4[k].x * m4[k].x + m4[k].y * m4[k].y + m4[k].z * m4[k].z + m4[k].w * m4[k].w;
  }
  ss = warp_sum(ss);
  const float am = al / fmaxf(sqrtf(ss), 1e-12f);          // a_j / |mem_j|  (F.normalize, criterion.py:720)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint2 wq = reinterpret_cast<const uint2*>(what + row * MH_D)[lane + 32 * k];
    const float2 w0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wq.x));
    const float2 w1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wq.y));
    __nv_bfloat162 p0 = __floats2bfloat162_rn(fmaf(be, w0.x, am * m4[k].x), fmaf(be, w0.y, am * m4[k].y));
    __nv_bfloat162 p1 = __floats2bfloat162_rn(fmaf(be, w1.x, am * m4[k].z), fmaf(be, w1.y, am * m4[k].w));
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&p0);
    pk.y = *reinterpret_cast<uint32_t*>(&p1);
    dst[lane + 32 * k] = pk;
  }
}

extern "C" int mh_vpl_mix(const void* w_hat_bf16, const float* mem, const float* life, float lamda, int64_t C, int64_t C_pad,
                          void* v_bf16, float* alpha_out, void* stream) {
  MH_CHECK_ARG(w_hat_bf16 && mem && life && v_bf16 && alpha_out, "null pointer");
  MH_CHECK_ARG(C > 0 && C_pad >= C, "bad shape");
  MH_CHECK_ARG(((uintptr_t)mem & 15) == 0 && ((uintptr_t)w_hat_bf16 & 7) == 0 && ((uintptr_t)v_bf16 & 7) == 0, "alignment");
  mh_launch(vpl_mix_kernel, (unsigned)((C_pad + 7) / 8), 256, 0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)w_hat_bf16, mem, life, lamda, C, C_pad, (__nv_bfloat16*)v_bf16, alpha_out);
  MH_LAUNCH_OK();
  return MH_OK;
}

// ------------------------------------------------------------------------------------------------
// prologue_x: one warp per embedding row. Reads x once (4*512 B in fp32), gathers the target class
// centre (4*512 B), writes x_hat bf16 (1 KB) + fp32 copy + three scalars.
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float load_as_float(const T* p, int64_t i);
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p, int64_t i) { return p[i]; }
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) { return __bfloat162float(p[i]); }
template <>
__device__ __forceinline__ float load_as_float<__half>(const __half* p, int64_t i) { return __half2float(p[i]); }

template <typename T>
__global__ void __launch_bounds__(256) prologue_x_kernel(const T* __restrict__ x, int64_t B, int64_t B_pad,
                                                         const int64_t* __restrict__ labels, const float* __restrict__ W,
                                                         int layout, int64_t C, int64_t ld, int64_t c_offset,
                                                         const float* __restrict__ inv_norm,
                                                         __nv_bfloat16* __restrict__ xhat, float* __restrict__ xhat32,
                                                         float* __restrict__ xnorm, float* __restrict__ t_raw,
                                                         int32_t* __restrict__ label_local, int64_t c_total) {
  mh_pdl_sync();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= B_pad) return;
  if (row >= B) {
    for (int k = lane; k < MH_D / 2; k += 32) reinterpret_cast<uint32_t*>(xhat + row * MH_D)[k] = 0u;
    if (lane == 0) label_local[row] = -1;
    return;
  }
  float v[16];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float f = load_as_float<T>(x, row * MH_D + (lane + 32 * k) * 4 + e);
      v[k * 4 + e] = f;
      ss += f * f;
    }
  }
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);
  const float inv = 1.f / fmaxf(nrm, 1e-12f);
  const int64_t y = labels[row] - c_offset;
  const bool owned = (y >= 0 && y < C);
  float dot = 0.f, wss = 0.f;       // wss: |w_y|^2 of the gathered row, same summation order as prologue_w_cd_kernel
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float4 o = make_float4(v[k * 4 + 0] * inv, v[k * 4 + 1] * inv, v[k * 4 + 2] * inv, v[k * 4 + 3] * inv);
    const int d = (lane + 32 * k) * 4;
    __nv_bfloat162 p0 = __floats2bfloat162_rn(o.x, o.y), p1 = __floats2bfloat162_rn(o.z, o.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&p0);
    pk.y = *reinterpret_cast<uint32_t*>(&p1);
    reinterpret_cast<uint2*>(xhat + row * MH_D)[lane + 32 * k] = pk;
    reinterpret_cast<float4*>(xhat32 + row * MH_D)[lane + 32 * k] = o;
    if (owned) {
      if (layout == MH_LAYOUT_CD) {
        float4 w = __ldg(reinterpret_cast<const float4*>(W + y * ld + d));
        dot += o.x * w.x + o.y * w.y + o.z * w.z + o.w * w.w;
        wss += w.x * w.x + w.y * w.y + w.z * w.z + w.w * w.w;
      } else {
        dot += o.x * __ldg(W + (int64_t)(d + 0) * ld + y) + o.y * __ldg(W + (int64_t)(d + 1) * ld + y) +
               o.z * __ldg(W + (int64_t)(d + 2) * ld + y) + o.w * __ldg(W + (int64_t)(d + 3) * ld + y);
      }
    }
  }
  dot = warp_sum(dot);
  // inv_norm == NULL (merged prologue + forward: 1/|w| is not written yet): take the norm of the gathered row here
  float inv_w = 0.f;
  if (inv_norm == nullptr) {
    wss = warp_sum(wss);
    inv_w = 1.f / fmaxf(sqrtf(wss), 1e-12f);
  }
  if (lane == 0) {
    xnorm[row] = nrm;
    // a label outside [0, c_total) (the whole head, all shards) poisons the row's target cosine, so the loss comes out
    // NaN instead of silently dropping the target (the reference's one_hot.scatter_ raises a device assert there);
    // sharded: a valid label owned by another rank contributes 0 to the all-reduce(SUM) of t_raw, NaN survives it
    const bool valid = labels[row] >= 0 && labels[row] < c_total;
    t_raw[row] = owned ? dot * (inv_norm ? inv_norm[y] : inv_w) : (valid ? 0.f : __int_as_float(0x7fc00000));
    label_local[row] = owned ? (int32_t)y : -1;
  }
}

extern "C" int mh_prologue_x(const void* x, int x_dtype, int64_t B, int64_t B_pad, const int64_t* labels,
                             const float* W, int layout, int64_t C, int64_t ld, int64_t c_offset,
                             const float* inv_norm, void* x_hat_bf16, float* x_hat32, float* xnorm, float* t_raw,
                             int32_t* label_local, int64_t c_total, void* stream) {
  MH_CHECK_ARG(x && labels && W && x_hat_bf16 && x_hat32 && xnorm && t_raw && label_local, "null pointer");
  MH_CHECK_ARG(inv_norm || layout == MH_LAYOUT_CD, "inv_norm may be NULL (norm taken from the gathered row) for the [C, 512] layout only");
  MH_CHECK_ARG(B > 0 && B_pad >= B && B_pad % MH_TILE == 0, "B_pad must be a multiple of 128 and >= B");
  MH_CHECK_ARG(layout == MH_LAYOUT_CD || layout == MH_LAYOUT_DC, "unknown layout");
  MH_CHECK_ARG(layout != MH_LAYOUT_CD || (ld % 4 == 0 && ((uintptr_t)W & 15) == 0), "CD weight must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((B_pad + 7) / 8));
  __nv_bfloat16* xh = (__nv_bfloat16*)x_hat_bf16;
  if (x_dtype == MH_F32)
    mh_launch(prologue_x_kernel<float>, grid, 256, 0, st, (const float*)x, B, B_pad, labels, W, layout, C, ld, c_offset,
                                                   inv_norm, xh, x_hat32, xnorm, t_raw, label_local, c_total);
  else if (x_dtype == MH_BF16)
    mh_launch(prologue_x_kernel<__nv_bfloat16>, grid, 256, 0, st, (const __nv_bfloat16*)x, B, B_pad, labels, W, layout, C, ld,
                                                           c_offset, inv_norm, xh, x_hat32, xnorm, t_raw, label_local, c_total);
  else if (x_dtype == MH_F16)
    mh_launch(prologue_x_kernel<__half>, grid, 256, 0, st, (const __half*)x, B, B_pad, labels, W, layout, C, ld, c_offset,
                                                    inv_norm, xh, x_hat32, xnorm, t_raw, label_local, c_total);
  else
    MH_CHECK_ARG(false, "unknown x dtype");
  MH_LAUNCH_OK();
  return MH_OK;
}

// ------------------------------------------------------------------------------------------------
// row_params: per-row margin terms and batch-global state. One block; B is a few thousand at most.
// Mirrors the scalar/vector parts of every reference forward (file:line beside each branch).
// ------------------------------------------------------------------------------------------------
__device__ double block_sum_dd(double v, double* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double t = 0.0;
  for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sh[k];
  return t;
}

__device__ __forceinline__ float cheb(int m, float c) {  // criterion.py:40-47
  switch (m) {
    case 0: return 1.f;
    case 1: return c;
    case 2: return 2.f * c * c - 1.f;
    case 3: return (4.f * c * c - 3.f) * c;
    case 4: return (8.f * c * c - 8.f) * c * c + 1.f;
    default: return ((16.f * c * c - 20.f) * c * c + 5.f) * c;
  }
}
__device__ __forceinline__ float dcheb(int m, float c) {
  switch (m) {
    case 0: return 0.f;
    case 1: return 1.f;
    case 2: return 4.f * c;
    case 3: return 12.f * c * c - 3.f;
    case 4: return (32.f * c * c - 16.f) * c;
    default: return (80.f * c * c - 60.f) * c * c + 5.f;
  }
}

__global__ void __launch_bounds__(1024) row_params_kernel(MhParams p, int64_t B, const float* __restrict__ xnorm,
                                                          const float* __restrict__ t_raw,
                                                          const float* __restrict__ margins, float* state,
                                                          int update_state, float* __restrict__ rowp, int64_t ldp) {
  mh_pdl_sync();
  __shared__ double sh[32];
  __shared__ float bc[4];
  const float PI = 3.14159265358979323846f;
  // ---- batch-global quantities ------------------------------------------------------------------
  if (p.family == MH_CURRICULAR) {
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < B; i += blockDim.x) acc += (double)fminf(fmaxf(t_raw[i], p.lo), p.hi);
    double tot = block_sum_dd(acc, sh);
    if (threadIdx.x == 0) {                       // criterion.py:570-573
      float t_new = (float)(tot / (double)B) * p.momentum + (1.f - p.momentum) * state[0];
      if (update_state) state[0] = t_new;
      state[4] = t_new;                           // the value criterion.py:575 multiplies with
      bc[0] = t_new;
    }
  } else if (p.family == MH_ADAFACE) {
    double a = 0.0;
    for (int64_t i = threadIdx.x; i < B; i += blockDim.x) a += (double)fminf(fmaxf(xnorm[i], 0.001f), 100.f);
    double mean = block_sum_dd(a, sh) / (double)B;
    double q = 0.0;
    for (int64_t i = threadIdx.x; i < B; i += blockDim.x) {
      double d = (double)fminf(fmaxf(xnorm[i], 0.001f), 100.f) - mean;
      q += d * d;
    }
    double var = block_sum_dd(q, sh) / (double)(B - 1);   // torch.std is unbiased (criterion.py:880)
    if (threadIdx.x == 0) {                               // criterion.py:881-882
      float bm = (float)mean * p.t_alpha + (1.f - p.t_alpha) * state[1];
      float bs = (float)sqrt(var) * p.t_alpha + (1.f - p.t_alpha) * state[2];
      if (update_state) { state[1] = bm; state[2] = bs; }
      bc[1] = bm; bc[2] = bs;
    }
  } else if (p.family == MH_MAGFACE) {
    double a = 0.0;
    for (int64_t i = threadIdx.x; i < B; i += blockDim.x) {
      float xc = fminf(fmaxf(xnorm[i], p.l_a), p.u_a);
      a += (double)(xc / (p.u_a * p.u_a) + 1.f / xc);     // criterion.py:1237
    }
    double tot = block_sum_dd(a, sh);
    if (threadIdx.x == 0) state[3] = (float)(tot / (double)B);
  } else {
    if (threadIdx.x == 0) state[3] = 0.f;
  }
  __syncthreads();
  // ---- per-row terms ----------------------------------------------------------------------------
  for (int64_t i = threadIdx.x; i < B; i += blockDim.x) {
    const float xn = xnorm[i];
    const float tr = t_raw[i];
    const float t = fminf(fmaxf(tr, p.lo), p.hi);
    const float inside = (tr >= p.lo && tr <= p.hi) ? 1.f : 0.f;
    float scale = p.s, thr = INFINITY, zt = 0.f, dzt = 0.f, dzt_dn = 0.f, dlg_dn = 0.f, norms = xn;
    switch (p.family) {
      case MH_ARCFACE: {                                  // criterion.py:281-287
        float one_m = 1.f - t * t;
        float sine = sqrtf(fminf(fmaxf(one_m, 1e-9f), 1.f));
        float dsine = (one_m >= 1e-9f && one_m <= 1.f) ? -t / sine : 0.f;
        float phi = t * p.cos_m - sine * p.sin_m;
        float dphi = p.cos_m - dsine * p.sin_m;
        bool take = p.easy_margin ? (t > 0.f) : (t > p.th);
        float alt = p.easy_margin ? t : t - p.mm;
        zt = p.s * (take ? phi : alt);
        dzt = p.s * (take ? dphi : 1.f);
      } break;
      case MH_COSFACE:                                    // criterion.py:186-189
        zt = p.s * (t - p.m);
        dzt = p.s * inside;
        break;
      case MH_SPHEREFACE: {                               // criterion.py:85-105
        float theta = acosf(t);
        float k = floorf((float)p.sphere_m * theta / PI);
        float sign = (fmodf(k, 2.f) == 0.f) ? 1.f : -1.f;
        float phi = sign * cheb(p.sphere_m, t) - 2.f * k;
        float u = (phi - t) / (1.f + p.sphere_lambda) + t;
        float du = (sign * dcheb(p.sphere_m, t) - 1.f) / (1.f + p.sphere_lambda) + 1.f;
        scale = xn;
        zt = u * xn;
        dzt = du * xn * inside;
      } break;
      case MH_MV_AM: {                                    // criterion.py:421-424
        bool take = t > p.m;
        zt = p.s * (take ? t - p.m : t);
        dzt = p.s * inside;
        thr = t - p.m;
      } break;
      case MH_MV_ARC: {                                   // criterion.py:427-430
        float sin_t = sqrtf(1.f - t * t + 1e-9f);
        float ctm = t * p.cos_m - sin_t * p.sin_m;
        bool take = t > 0.f;
        zt = p.s * (take ? ctm : t);
        dzt = p.s * (take ? p.cos_m + t / sin_t * p.sin_m : 1.f) * inside;
        thr = ctm;
      } break;
      case MH_CURRICULAR: {                               // criterion.py:555-566
        float sin_t = sqrtf(1.f - t * t);
        float ctm = t * p.cos_m - sin_t * p.sin_m;
        bool take = t > p.th;
        zt = p.s * (take ? ctm : t - p.mm);
        dzt = p.s * (take ? p.cos_m + t / sin_t * p.sin_m : 1.f) * inside;
        thr = ctm;
      } break;
      case MH_ADAFACE: {                                  // criterion.py:876-904
        const float eps = 1e-3f;
        float sn = fminf(fmaxf(xn, 0.001f), 100.f);
        float ms = fminf(fmaxf((sn - bc[1]) / (bc[2] + eps) * p.h, -1.f), 1.f);
        float th_raw = acosf(t) - p.m * ms;
        float th_m = fminf(fmaxf(th_raw, eps), PI - eps);
        float in2 = (th_raw >= eps && th_raw <= PI - eps) ? 1.f : 0.f;
        zt = p.s * (cosf(th_m) - (p.m + p.m * ms));
        dzt = p.s * sinf(th_m) / sqrtf(1.f - t * t) * in2 * inside;
      } break;
      case MH_ELASTIC_COS: {                              // criterion.py:1014
        zt = p.s * (t - margins[i]);
        dzt = p.s * inside;
      } break;
      case MH_ELASTIC_ARC: {                              // criterion.py:1129-1135
        float th_raw = acosf(t) + margins[i];
        float th_m = fminf(fmaxf(th_raw, 0.f), PI);
        float in2 = (th_raw >= 0.f && th_raw <= PI) ? 1.f : 0.f;
        zt = p.s * cosf(th_m);
        dzt = p.s * sinf(th_m) / sqrtf(1.f - t * t) * in2 * inside;
      } break;
      case MH_VPL_ARC: {                                  // criterion.py:724-749 (target column: cosine2 -> clamp -> margin)
        // margins[i] = a_y = lamda * 1[life_y > 0]; target cosine = (1 - a_y) * <x^, w^_y> + a_y, both weights in fp32
        const float al = margins[i], be = 1.f - al;
        const float c2 = fmaf(be, tr, al);
        const float tc = fminf(fmaxf(c2, p.lo), p.hi);
        const float in2 = (c2 >= p.lo && c2 <= p.hi) ? 1.f : 0.f;
        const float sin_t = sqrtf(1.f - tc * tc + 1e-9f);
        const float phi = tc * p.cos_m - sin_t * p.sin_m;
        const float dphi = p.cos_m + tc / sin_t * p.sin_m;
        const bool take = p.easy_margin ? (tc > 0.f) : (tc > p.th);
        const float alt = p.easy_margin ? tc : tc - p.mm;
        zt = p.s * (take ? phi : alt);
        dzt = p.s * (take ? dphi : 1.f) * in2 * be;       // d zt / d <x^, w^_y>
        rowp[MH_RP_T * ldp + i] = tc;                     // pre-margin value at the target column (rank count)
      } break;
      case MH_MAGFACE: {                                  // criterion.py:1244-1278
        float xc = fminf(fmaxf(xn, p.l_a), p.u_a);
        float in_n = (xn >= p.l_a && xn <= p.u_a) ? 1.f : 0.f;
        float ks = (p.u_margin - p.l_margin) / (p.u_a - p.l_a);
        float a = ks * (xc - p.l_a) + p.l_margin;
        float ca = cosf(a), sa = sinf(a);
        float sin_t = sqrtf(1.f - t * t + 1e-9f);
        float ctm = t * ca - sin_t * sa;
        float dctm_dt = ca + t / sin_t * sa;
        float dctm_da = -t * sa - sin_t * ca;
        bool take;
        float alt, dalt_da;
        if (p.easy_margin) { take = t > 0.f; alt = t; dalt_da = 0.f; }
        else { take = t > cosf(PI - a); alt = t - sinf(PI - a) * a; dalt_da = -(a * ca + sa); }
        zt = p.s * (take ? ctm : alt);
        dzt = p.s * (take ? dctm_dt : 1.f) * inside;
        dzt_dn = p.s * (take ? dctm_da : dalt_da) * ks * in_n;
        dlg_dn = (1.f / (p.u_a * p.u_a) - 1.f / (xc * xc)) / (float)B * in_n;
        norms = xc;
      } break;
    }
    if (tr != tr) zt = tr;              // poisoned target cosine (label out of range): keep the NaN visible in the loss
    rowp[MH_RP_SCALE * ldp + i] = scale;
    rowp[MH_RP_THR * ldp + i] = thr;
    rowp[MH_RP_ZT * ldp + i] = zt;
    rowp[MH_RP_DZT * ldp + i] = dzt;
    if (p.family != MH_VPL_ARC) rowp[MH_RP_T * ldp + i] = t;
    rowp[MH_RP_DZT_DN * ldp + i] = dzt_dn;
    rowp[MH_RP_DLG_DN * ldp + i] = dlg_dn;
    rowp[MH_RP_NORMS * ldp + i] = norms;
  }
  // padding rows: benign values (scale 0 -> z = 0)
  for (int64_t i = B + threadIdx.x; i < ldp; i += blockDim.x) {
#pragma unroll
    for (int k = 0; k < MH_RP_PLANES; ++k) rowp[k * ldp + i] = (k == MH_RP_THR) ? INFINITY : 0.f;
  }
}

extern "C" int mh_row_params(const mh_config* cfg_host, int64_t B, const float* xnorm, const float* t_raw,
                             const float* margins, float* state, int update_state, float* rowp, int64_t ldp,
                             void* stream) {
  MH_CHECK_ARG(cfg_host && xnorm && t_raw && state && rowp, "null pointer");
  MH_CHECK_ARG(B > 0 && ldp >= B, "ldp must be >= B");
  MhParams p = mh_make_params(cfg_host);
  MH_CHECK_ARG(p.family >= 0 && p.family <= MH_VPL_ARC, "unknown family");
  MH_CHECK_ARG((p.family != MH_ELASTIC_COS && p.family != MH_ELASTIC_ARC && p.family != MH_VPL_ARC) || margins,
               "ElasticFace / VPL-ArcFace need the per-row margins / interpolation weights");
  MH_CHECK_ARG(p.family != MH_SPHEREFACE || (p.sphere_m >= 0 && p.sphere_m <= 5), "SphereFace m must be in 0..5");
  mh_launch(row_params_kernel, 1, 1024, 0, (cudaStream_t)stream, p, B, xnorm, t_raw, margins, state, update_state, rowp, ldp);
  MH_LAUNCH_OK();
  return MH_OK;
}
