// Exact fp32 path (SIMT) and the O(B)/O(C*d) reductions shared by both paths.
//
// The exact path materialises S = x^ w^T [B, C] in fp32 (small C: tests, the fp32-tolerance mode and
// the compat mode that returns the reference's 4-tuple).  It is NOT the performance path; the
// tensor-core path (tc_head.cu) never materialises logits.
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// Generic strided SGEMM: C[M,N] = A[M,K] . B[K,N], arbitrary (row, col) element strides for A and B.
// 64x64 tile, BK=16, 256 threads, 4x4 outputs per thread, fp32 FMA.
// ------------------------------------------------------------------------------------------------
#define SG_BM 64
#define SG_BN 64
#define SG_BK 16
__global__ void __launch_bounds__(256) sgemm_strided_kernel(int64_t M, int64_t N, int64_t K, const float* __restrict__ A,
                                                            int64_t a_rs, int64_t a_cs, const float* __restrict__ Bm,
                                                            int64_t b_rs, int64_t b_cs, float* __restrict__ Cm,
                                                            int64_t ldc) {
  mh_pdl_sync();
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Bs[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * SG_BM, n0 = (int64_t)blockIdx.x * SG_BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // loader mapping: pick the thread->element map so that the unit-stride dimension runs along tid
  const bool a_k_contig = (a_cs == 1);
  const bool b_n_contig = (b_cs == 1);
  for (int64_t k0 = 0; k0 < K; k0 += SG_BK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int e = tid + r * 256;               // 0..1023 = 64 x 16
      int mm, kk;
      if (a_k_contig) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
      int64_t gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < K) ? A[gm * a_rs + gk * a_cs] : 0.f;
      int nn, kb;
      if (b_n_contig) { nn = e & 63; kb = e >> 6; } else { kb = e & 15; nn = e >> 4; }
      int64_t gn = n0 + nn, gkb = k0 + kb;
      Bs[kb][nn] = (gn < N && gkb < K) ? Bm[gkb * b_rs + gn * b_cs] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t gn = n0 + tx * 4 + j;
      if (gn < N) Cm[gm * ldc + gn] = acc[i][j];
    }
  }
}

extern "C" int mh_sgemm_strided(int64_t M, int64_t N, int64_t K, const float* A, int64_t a_rs, int64_t a_cs,
                                const float* B, int64_t b_rs, int64_t b_cs, float* C, int64_t ldc, void* stream) {
  MH_CHECK_ARG(A && B && C, "null pointer");
  MH_CHECK_ARG(M > 0 && N > 0 && K > 0 && ldc >= N, "bad shape");
  dim3 grid((unsigned)((N + SG_BN - 1) / SG_BN), (unsigned)((M + SG_BM - 1) / SG_BM));
  MH_CHECK_ARG(grid.y <= 65535, "M too large for the exact path");
  mh_launch(sgemm_strided_kernel, grid, 256, 0, (cudaStream_t)stream, M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, C, ldc);
  MH_LAUNCH_OK();
  return MH_OK;
}

// ------------------------------------------------------------------------------------------------
// dense_forward: one block per row over the materialised cosines.
// ------------------------------------------------------------------------------------------------
struct RowStat { float m, l, cnt, ez; };
__device__ __forceinline__ void stat_push(RowStat& s, float z2, float u) {
  if (z2 > s.m) {
    float r = exp2f(s.m - z2);
    s.l *= r; s.ez *= r; s.m = z2;
  }
  float e = exp2f(z2 - s.m);
  s.l += e;
  s.ez = fmaf(e, u, s.ez);
}
__device__ __forceinline__ void stat_merge(RowStat& a, const RowStat& b) {
  float m = fmaxf(a.m, b.m);
  float ra = (a.m == -INFINITY) ? 0.f : exp2f(a.m - m);
  float rb = (b.m == -INFINITY) ? 0.f : exp2f(b.m - m);
  a.l = a.l * ra + b.l * rb;
  a.ez = a.ez * ra + b.ez * rb;
  a.cnt += b.cnt;
  a.m = m;
}

__global__ void __launch_bounds__(256) dense_forward_kernel(MhParams p, const float* __restrict__ S, int64_t lds_,
                                                            int64_t B, int64_t B_pad, int64_t C,
                                                            const float* __restrict__ rowp, int64_t ldp,
                                                            const int32_t* __restrict__ label_local,
                                                            const float* __restrict__ state, float* __restrict__ stats,
                                                            float* __restrict__ pre, float* __restrict__ logits) {
  mh_pdl_sync();
  __shared__ RowStat sh[8];
  const int64_t i = blockIdx.x;
  const float scale = rowp[MH_RP_SCALE * ldp + i], thr = rowp[MH_RP_THR * ldp + i];
  const float zt = rowp[MH_RP_ZT * ldp + i], t = rowp[MH_RP_T * ldp + i];
  const int32_t y = label_local[i];
  const float ha = (p.hard_kind == 2) ? state[4] : p.hard_a;
  RowStat st{-INFINITY, 0.f, 0.f, 0.f};
  for (int64_t j = threadIdx.x; j < C; j += blockDim.x) {
    ElemOut e = mh_elem(S[i * lds_ + j], p.lo, p.hi, p.hard_kind, thr, ha, p.hard_b);
    float z = scale * e.u, u = e.u;
    if (j == y) { z = zt; u = zt / scale; }
    else if (e.c > t) st.cnt += 1.f;
    stat_push(st, z * MH_LOG2E, u);
    if (pre) pre[i * C + j] = scale * e.c;
    if (logits) logits[i * C + j] = z;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    RowStat b;
    b.m = __shfl_xor_sync(0xffffffffu, st.m, o);
    b.l = __shfl_xor_sync(0xffffffffu, st.l, o);
    b.cnt = __shfl_xor_sync(0xffffffffu, st.cnt, o);
    b.ez = __shfl_xor_sync(0xffffffffu, st.ez, o);
    stat_merge(st, b);
  }
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = st;
  __syncthreads();
  if (threadIdx.x == 0) {
    RowStat a = sh[0];
    for (int k = 1; k < 8; ++k) stat_merge(a, sh[k]);
    stats[MH_ST_M * B_pad + i] = a.m;
    stats[MH_ST_L * B_pad + i] = a.l;
    stats[MH_ST_CNT * B_pad + i] = a.cnt;
    stats[MH_ST_EZ * B_pad + i] = a.ez;
  }
}

extern "C" int mh_dense_forward(const mh_config* cfg_host, const float* S, int64_t lds_, int64_t B, int64_t B_pad,
                                int64_t C, const float* rowp, int64_t ldp, const int32_t* label_local,
                                const float* state, float* stats, float* pre, float* logits, void* stream) {
  MH_CHECK_ARG(cfg_host && S && rowp && label_local && state && stats, "null pointer");
  MH_CHECK_ARG(B > 0 && C > 0 && lds_ >= C && B_pad >= B, "bad shape");
  MhParams p = mh_make_params(cfg_host);
  mh_launch(dense_forward_kernel, (unsigned)B, 256, 0, (cudaStream_t)stream, p, S, lds_, B, B_pad, C, rowp, ldp, label_local,
                                                                    state, stats, pre, logits);
  MH_LAUNCH_OK();
  return MH_OK;
}

// ------------------------------------------------------------------------------------------------
// dense_backward_dc: S -> dcos in place.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dense_backward_dc_kernel(MhParams p, float* __restrict__ S, int64_t lds_,
                                                                int64_t B, int64_t C, const float* __restrict__ rowp,
                                                                int64_t ldp, const int32_t* __restrict__ label_local,
                                                                const float* __restrict__ state,
                                                                const float* __restrict__ lse2,
                                                                const float* __restrict__ dlogits,
                                                                const float* __restrict__ dpre, float* __restrict__ rowaux) {
  mh_pdl_sync();
  __shared__ float red[8];
  const int64_t i = blockIdx.x;
  const float scale = rowp[MH_RP_SCALE * ldp + i], thr = rowp[MH_RP_THR * ldp + i];
  const float zt = rowp[MH_RP_ZT * ldp + i], dzt = rowp[MH_RP_DZT * ldp + i];
  const int32_t y = label_local[i];
  const float ha = (p.hard_kind == 2) ? state[4] : p.hard_a;
  const float l2 = lse2 ? lse2[i] : 0.f;
  float aux1 = 0.f;
  for (int64_t j = threadIdx.x; j < C; j += blockDim.x) {
    const float raw = S[i * lds_ + j];
    ElemOut e = mh_elem(raw, p.lo, p.hi, p.hard_kind, thr, ha, p.hard_b);
    const float inside = (raw >= p.lo && raw <= p.hi) ? 1.f : 0.f;
    float z = scale * e.u, dzdc = scale * e.du, u = e.u;
    if (j == y) { z = zt; dzdc = dzt; u = zt / scale; }
    float dz, dc;
    if (dlogits) {
      dz = dlogits[i * C + j];
      dc = dz * dzdc;
      aux1 += dz * u;
      if (dpre) {
        float dp = dpre[i * C + j];
        dc += dp * scale * inside;
        aux1 += dp * e.c;
      }
      if (j == y && rowaux) rowaux[i] = dz;
    } else {
      dz = exp2f(z * MH_LOG2E - l2) - (j == y ? 1.f : 0.f);
      dc = dz * dzdc;
    }
    S[i * lds_ + j] = dc;
  }
  if (dlogits && rowaux) {
    aux1 = warp_sum(aux1);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = aux1;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int k = 0; k < 8; ++k) t += red[k];
      rowaux[B + i] = p.scale_is_norm ? t : 0.f;
      if (y < 0) rowaux[i] = 0.f;
    }
  }
}

extern "C" int mh_dense_backward_dc(const mh_config* cfg_host, float* S, int64_t lds_, int64_t B, int64_t C,
                                    const float* rowp, int64_t ldp, const int32_t* label_local, const float* state,
                                    const float* lse2, const float* dlogits, const float* dpre, float* rowaux,
                                    void* stream) {
  MH_CHECK_ARG(cfg_host && S && rowp && label_local && state, "null pointer");
  MH_CHECK_ARG((lse2 != nullptr) != (dlogits != nullptr), "exactly one of lse2 / dlogits must be given");
  MH_CHECK_ARG(!dlogits || rowaux, "compat mode needs rowaux");
  MhParams p = mh_make_params(cfg_host);
  mh_launch(dense_backward_dc_kernel, (unsigned)B, 256, 0, (cudaStream_t)stream, p, S, lds_, B, C, rowp, ldp, label_local,
                                                                        state, lse2, dlogits, dpre, rowaux);
  MH_LAUNCH_OK();
  return MH_OK;
}

// ------------------------------------------------------------------------------------------------
// merge_stats: online-softmax merge over tiles / shards.  grid (rows/128, nblk); block (128, 8).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) merge_stats_kernel(const float* __restrict__ in, int64_t n_parts, int64_t lds_,
                                                           float* __restrict__ out, const int* gate, int gate_on) {
  mh_pdl_sync();
  __shared__ RowStat sh[8][128];
  if (gate && ((*reinterpret_cast<const volatile int*>(gate) != 0) != (gate_on != 0))) return;   // guarded stash, see mh_step_forward
  const int rx = threadIdx.x, py = threadIdx.y;
  const int64_t row = (int64_t)blockIdx.x * 128 + rx;
  const int64_t per = (n_parts + gridDim.y - 1) / gridDim.y;
  const int64_t p0 = (int64_t)blockIdx.y * per, p1 = min(n_parts, p0 + per);
  RowStat a{-INFINITY, 0.f, 0.f, 0.f};
  if (row < lds_) {
    for (int64_t q = p0 + py; q < p1; q += 8) {
      const float* b = in + q * MH_ST_PLANES * lds_;
      RowStat s{b[MH_ST_M * lds_ + row], b[MH_ST_L * lds_ + row], b[MH_ST_CNT * lds_ + row], b[MH_ST_EZ * lds_ + row]};
      stat_merge(a, s);
    }
  }
  sh[py][rx] = a;
  __syncthreads();
  if (py == 0 && row < lds_) {
    for (int k = 1; k < 8; ++k) stat_merge(a, sh[k][rx]);
    float* o = out + (int64_t)blockIdx.y * MH_ST_PLANES * lds_;
    o[MH_ST_M * lds_ + row] = a.m;
    o[MH_ST_L * lds_ + row] = a.l;
    o[MH_ST_CNT * lds_ + row] = a.cnt;
    o[MH_ST_EZ * lds_ + row] = a.ez;
  }
}

extern "C" int mh_merge_stats(const float* stats_in, int64_t n_parts, int64_t B, int64_t lds_, float* scratch,
                              float* stats_out, void* stream) {
  return mh_merge_stats_ex(stats_in, n_parts, B, lds_, scratch, stats_out, nullptr, 0, stream);
}

extern "C" int mh_merge_stats_ex(const float* stats_in, int64_t n_parts, int64_t B, int64_t lds_, float* scratch, float* stats_out,
                        const int* gate, int gate_on, void* stream) {
  MH_CHECK_ARG(stats_in && stats_out && n_parts > 0 && lds_ >= B, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 block(128, 8);
  unsigned gx = (unsigned)((lds_ + 127) / 128);
  int64_t nblk = (n_parts >= 256) ? MH_MERGE_BLOCKS : 1;
  if (nblk > 1) {
    MH_CHECK_ARG(scratch, "scratch required for large merges");
    mh_launch(merge_stats_kernel, dim3(gx, (unsigned)nblk), block, 0, st, stats_in, n_parts, lds_, scratch, gate, gate_on);
    mh_launch(merge_stats_kernel, dim3(gx, 1), block, 0, st, scratch, nblk, lds_, stats_out, gate, gate_on);
  } else {
    mh_launch(merge_stats_kernel, dim3(gx, 1), block, 0, st, stats_in, n_parts, lds_, stats_out, gate, gate_on);
  }
  MH_LAUNCH_OK();
  return MH_OK;
}

// ------------------------------------------------------------------------------------------------
// finalize_rows: lse, per-row loss, rank counts, AUX planes and the batch scalars.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) finalize_rows_kernel(const float* __restrict__ stats, int64_t lds_,
                                                             const float* __restrict__ rowp, int64_t ldp, int64_t B,
                                                             int64_t B_total, int sphere, float* __restrict__ rowout,
                                                             int64_t ldo, float* __restrict__ scalars,
                                                             const float* __restrict__ state, float guard_min_l,
                                                             int* __restrict__ guard_flag, const int* gate, int gate_on) {
  mh_pdl_sync();
  __shared__ double sh[3][32];
  if (gate && ((*reinterpret_cast<const volatile int*>(gate) != 0) != (gate_on != 0))) return;   // guarded stash, see mh_step_forward
  double sl = 0.0, s1 = 0.0, s5 = 0.0;
  int unsafe = 0;
  for (int64_t i = threadIdx.x; i < B; i += blockDim.x) {
    const float m = stats[MH_ST_M * lds_ + i], l = stats[MH_ST_L * lds_ + i];
    // guarded stash: the fixed-reference sums of this row are exact to 2^-24 only if they dwarf everything that can have
    // been flushed to zero (mh_tc_stash_guarded_ok); written as !(l >= min) so that a NaN row is not "safe" by accident
    if (guard_flag && !(l >= guard_min_l) && l == l) unsafe = 1;
    const float cnt = stats[MH_ST_CNT * lds_ + i], ez = stats[MH_ST_EZ * lds_ + i];
    const float zt = rowp[MH_RP_ZT * ldp + i], scale = rowp[MH_RP_SCALE * ldp + i];
    const float lse2 = m + log2f(l);
    const float loss = lse2 * MH_LN2 - zt;
    const float pt = exp2f(zt * MH_LOG2E - lse2);
    rowout[MH_RO_LSE2 * ldo + i] = lse2;
    rowout[MH_RO_LOSS * ldo + i] = loss;
    rowout[MH_RO_CNT * ldo + i] = cnt;
    rowout[MH_RO_AUX0 * ldo + i] = pt - 1.f;
    rowout[MH_RO_AUX1 * ldo + i] = sphere ? (ez / l - zt / scale) : 0.f;
    sl += (double)loss;
    s1 += (cnt < 1.f) ? 1.0 : 0.0;
    s5 += (cnt < 5.f) ? 1.0 : 0.0;
  }
  double v[3] = {sl, s1, s5};
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (lane == 0) sh[k][w] = v[k];
  }
  if (guard_flag) {
    unsafe = __syncthreads_or(unsafe);
    if (threadIdx.x == 0) *guard_flag = unsafe ? 1 : 0;
  } else {
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double t[3] = {0, 0, 0};
    for (int k = 0; k < 3; ++k)
      for (int q = 0; q < 32; ++q) t[k] += sh[k][q];
    scalars[0] = (float)(t[0] / (double)B_total);
    scalars[1] = (float)(100.0 * t[1] / (double)B_total);
    scalars[2] = (float)(100.0 * t[2] / (double)B_total);
    scalars[3] = state ? state[3] : 0.f;                 // loss_g of this forward (MagFace), else 0
  }
}

extern "C" int mh_finalize_rows(const float* stats, int64_t lds_, const float* rowp, int64_t ldp, int64_t B,
                                int64_t B_total, int sphere, float* rowout, int64_t ldo, float* scalars,
                                const float* state, void* stream) {
  return mh_finalize_rows_ex(stats, lds_, rowp, ldp, B, B_total, sphere, rowout, ldo, scalars, state, 0.f, nullptr,
                               nullptr, 0, stream);
}

// guard_flag != NULL: also decide whether the fixed-reference sums can be trusted (every row sum >= guard_min_l) and write
// 0 / 1 to *guard_flag.  gate: the whole launch is a no-op unless (*gate != 0) == (gate_on != 0).
extern "C" int mh_finalize_rows_ex(const float* stats, int64_t lds_, const float* rowp, int64_t ldp, int64_t B, int64_t B_total,
                          int sphere, float* rowout, int64_t ldo, float* scalars, const float* state, float guard_min_l,
                          int* guard_flag, const int* gate, int gate_on, void* stream) {
  MH_CHECK_ARG(stats && rowp && rowout && scalars, "null pointer");
  MH_CHECK_ARG(B > 0 && B_total >= B && lds_ >= B && ldp >= B && ldo >= B, "bad shape");
  mh_launch(finalize_rows_kernel, 1, 1024, 0, (cudaStream_t)stream, stats, lds_, rowp, ldp, B, B_total, sphere, rowout, ldo,
                                                             scalars, state, guard_min_l, guard_flag, gate, gate_on);
  MH_LAUNCH_OK();
  return MH_OK;
}

// ------------------------------------------------------------------------------------------------
// norm_backward_x: one warp per row.
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void store_from_float4(T* p, float4 v);
template <>
__device__ __forceinline__ void store_from_float4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <>
__device__ __forceinline__ void store_from_float4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&a);
  pk.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = pk;
}
template <>
__device__ __forceinline__ void store_from_float4<__half>(__half* p, float4 v) {
  __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&a);
  pk.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = pk;
}

template <typename T>
__global__ void __launch_bounds__(256) norm_backward_x_kernel(const float* __restrict__ dxhat, int n_split,
                                                              int64_t split_stride, const float* __restrict__ xhat32,
                                                              const float* __restrict__ xnorm,
                                                              const float* __restrict__ rowp, int64_t ldp,
                                                              const float* __restrict__ aux0,
                                                              const float* __restrict__ aux1,
                                                              const float* __restrict__ gscal, int64_t B,
                                                              T* __restrict__ dx) {
  mh_pdl_sync();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= B) return;
  const float gz = gscal[0], glg = gscal[1];
  float4 g[4], xh[4];
  float dot = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < n_split; ++s) {
      float4 q = reinterpret_cast<const float4*>(dxhat + (int64_t)s * split_stride + row * MH_D)[lane + 32 * k];
      a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
    }
    a.x *= gz; a.y *= gz; a.z *= gz; a.w *= gz;
    g[k] = a;
    xh[k] = reinterpret_cast<const float4*>(xhat32 + row * MH_D)[lane + 32 * k];
    dot += a.x * xh[k].x + a.y * xh[k].y + a.z * xh[k].z + a.w * xh[k].w;
  }
  dot = warp_sum(dot);
  const float inv = 1.f / fmaxf(xnorm[row], 1e-12f);
  const float dn = gz * (aux0[row] * rowp[MH_RP_DZT_DN * ldp + row] + aux1[row]) + glg * rowp[MH_RP_DLG_DN * ldp + row];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float4 o;
    o.x = (g[k].x - xh[k].x * dot) * inv + dn * xh[k].x;
    o.y = (g[k].y - xh[k].y * dot) * inv + dn * xh[k].y;
    o.z = (g[k].z - xh[k].z * dot) * inv + dn * xh[k].z;
    o.w = (g[k].w - xh[k].w * dot) * inv + dn * xh[k].w;
    store_from_float4<T>(dx + row * MH_D + (lane + 32 * k) * 4, o);
  }
}

extern "C" int mh_norm_backward_x(const float* dxhat, int n_split, int64_t split_stride, const float* x_hat32,
                                  const float* xnorm, const float* rowp, int64_t ldp, const float* aux0,
                                  const float* aux1, const float* gscal, int64_t B, void* dx, int x_dtype,
                                  void* stream) {
  MH_CHECK_ARG(dxhat && x_hat32 && xnorm && rowp && aux0 && aux1 && gscal && dx, "null pointer");
  MH_CHECK_ARG(n_split >= 1 && B > 0, "bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((B + 7) / 8));
  if (x_dtype == MH_F32)
    mh_launch(norm_backward_x_kernel<float>, grid, 256, 0, st, dxhat, n_split, split_stride, x_hat32, xnorm, rowp, ldp, aux0,
                                                        aux1, gscal, B, (float*)dx);
  else if (x_dtype == MH_BF16)
    mh_launch(norm_backward_x_kernel<__nv_bfloat16>, grid, 256, 0, st, dxhat, n_split, split_stride, x_hat32, xnorm, rowp, ldp,
                                                                aux0, aux1, gscal, B, (__nv_bfloat16*)dx);
  else if (x_dtype == MH_F16)
    mh_launch(norm_backward_x_kernel<__half>, grid, 256, 0, st, dxhat, n_split, split_stride, x_hat32, xnorm, rowp, ldp, aux0,
                                                         aux1, gscal, B, (__half*)dx);
  else
    MH_CHECK_ARG(false, "unknown dtype");
  MH_LAUNCH_OK();
  return MH_OK;
}

// ------------------------------------------------------------------------------------------------
// norm_backward_w: dW_j = g (dw^_j - w^_j (w^_j.dw^_j)) / |w_j| in the parameter's own layout.
// CD: one warp per class, coalesced row writes.  DC: 32-class slab transposed through shared memory.
// Algorithmic bytes per class: 2048 (dw^) + 1024 (w^ bf16) read + 2048 written.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 load_what4(const __nv_bfloat16* wb, const float* w32, int64_t row, int idx4) {
  if (w32) return reinterpret_cast<const float4*>(w32 + row * MH_D)[idx4];
  uint2 pk = reinterpret_cast<const uint2*>(wb + row * MH_D)[idx4];
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&pk.x), b = *reinterpret_cast<__nv_bfloat162*>(&pk.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}

__global__ void __launch_bounds__(256) norm_backward_w_cd_kernel(const float* __restrict__ dwh,
                                                                 const __nv_bfloat16* __restrict__ wb,
                                                                 const float* __restrict__ w32,
                                                                 const float* __restrict__ inv_norm,
                                                                 const float* __restrict__ gscal,
                                                                 const float* __restrict__ class_scale, int64_t C,
                                                                 float* __restrict__ dW, int64_t ld) {
  mh_pdl_sync();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= C) return;
  float4 g[4], w[4];
  float dot = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    g[k] = reinterpret_cast<const float4*>(dwh + row * MH_D)[lane + 32 * k];
    w[k] = load_what4(wb, w32, row, lane + 32 * k);
    dot += g[k].x * w[k].x + g[k].y * w[k].y + g[k].z * w[k].z + g[k].w * w[k].w;
  }
  dot = warp_sum(dot);
  const float sc = gscal[0] * inv_norm[row] * (class_scale ? class_scale[row] : 1.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float4 o = make_float4((g[k].x - w[k].x * dot) * sc, (g[k].y - w[k].y * dot) * sc, (g[k].z - w[k].z * dot) * sc,
                           (g[k].w - w[k].w * dot) * sc);
    reinterpret_cast<float4*>(dW + row * ld)[lane + 32 * k] = o;
  }
}

__global__ void __launch_bounds__(256) norm_backward_w_dc_kernel(const float* __restrict__ dwh,
                                                                 const __nv_bfloat16* __restrict__ wb,
                                                                 const float* __restrict__ w32,
                                                                 const float* __restrict__ inv_norm,
                                                                 const float* __restrict__ gscal,
                                                                 const float* __restrict__ class_scale, int64_t C,
                                                                 float* __restrict__ dW, int64_t ld) {
  mh_pdl_sync();
  extern __shared__ float slab[];            // [512][33]
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c0 = (int64_t)blockIdx.x * 32;
  const float gz = gscal[0];
  for (int r = ty; r < 32; r += 8) {
    const int64_t row = c0 + r;
    float4 g[4], w[4];
    float dot = 0.f;
    if (row < C) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        g[k] = reinterpret_cast<const float4*>(dwh + row * MH_D)[tx + 32 * k];
        w[k] = load_what4(wb, w32, row, tx + 32 * k);
        dot += g[k].x * w[k].x + g[k].y * w[k].y + g[k].z * w[k].z + g[k].w * w[k].w;
      }
    }
    dot = warp_sum(dot);
    const float sc = (row < C) ? gz * inv_norm[row] * (class_scale ? class_scale[row] : 1.f) : 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int d = (tx + 32 * k) * 4;
      if (row < C) {
        slab[(d + 0) * 33 + r] = (g[k].x - w[k].x * dot) * sc;
        slab[(d + 1) * 33 + r] = (g[k].y - w[k].y * dot) * sc;
        slab[(d + 2) * 33 + r] = (g[k].z - w[k].z * dot) * sc;
        slab[(d + 3) * 33 + r] = (g[k].w - w[k].w * dot) * sc;
      }
    }
  }
  __syncthreads();
  const int64_t c = c0 + tx;
  if (c < C) {
    for (int d = ty; d < MH_D; d += 8) dW[(int64_t)d * ld + c] = slab[d * 33 + tx];
  }
}

extern "C" int mh_norm_backward_w(const float* dw_hat, const void* w_hat_bf16, const float* w_hat32,
                                  const float* inv_norm, const float* gscal, const float* class_scale, int64_t C,
                                  int layout, float* dW, int64_t ld, void* stream) {
  MH_CHECK_ARG(dw_hat && inv_norm && gscal && dW, "null pointer");
  MH_CHECK_ARG((w_hat_bf16 != nullptr) != (w_hat32 != nullptr), "exactly one of w_hat_bf16 / w_hat32");
  cudaStream_t st = (cudaStream_t)stream;
  if (layout == MH_LAYOUT_CD) {
    MH_CHECK_ARG(ld % 4 == 0 && ((uintptr_t)dW & 15) == 0, "dW must be 16-byte aligned");
    mh_launch(norm_backward_w_cd_kernel, (unsigned)((C + 7) / 8), 256, 0, st, dw_hat, (const __nv_bfloat16*)w_hat_bf16, w_hat32,
                                                                      inv_norm, gscal, class_scale, C, dW, ld);
  } else if (layout == MH_LAYOUT_DC) {
    static MhDeviceOnce attr_once;
    const int smem = MH_D * 33 * sizeof(float);
    MH_CUDA_OK(mh_once_per_device(attr_once, [&] {
      return cudaFuncSetAttribute(norm_backward_w_dc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    }));
    mh_launch(norm_backward_w_dc_kernel, (unsigned)((C + 31) / 32), 256, smem, st, dw_hat, (const __nv_bfloat16*)w_hat_bf16,
                                                                          w_hat32, inv_norm, gscal, class_scale, C, dW, ld);
  } else {
    MH_CHECK_ARG(false, "unknown layout");
  }
  MH_LAUNCH_OK();
  return MH_OK;
}

// ------------------------------------------------------------------------------------------------
// gscal = {upstream grad of loss_id / B_total, upstream grad of loss_g}: one tiny launch instead of a chain of
// framework fill / index kernels, and no host sync under a GradScaler (model_utils.py:185).
// ------------------------------------------------------------------------------------------------
__global__ void make_gscal_kernel(const float* __restrict__ g_loss, const float* __restrict__ g_lossg, float inv_b,
                                  float* __restrict__ gscal) {
  mh_pdl_sync();
  gscal[0] = (g_loss ? g_loss[0] : 0.f) * inv_b;
  gscal[1] = g_lossg ? g_lossg[0] : 0.f;
}

extern "C" int mh_make_gscal(const float* g_loss, const float* g_lossg, int64_t B_total, float* gscal, void* stream) {
  MH_CHECK_ARG(gscal && B_total > 0, "bad argument");
  mh_launch(make_gscal_kernel, 1, 1, 0, (cudaStream_t)stream, g_loss, g_lossg, 1.f / (float)B_total, gscal);
  MH_LAUNCH_OK();
  return MH_OK;
}
