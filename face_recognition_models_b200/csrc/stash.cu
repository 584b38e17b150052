// Small O(B*d) kernels of the stash backward: with the forward stash E'_ij = exp2(z_ij log2e - ref_i) * du_ij/dcos
// (target column = 0) the softmax-CE gradient factorises as
//     G_ij = rho_i * E'_ij            (j != y_i),   rho_i = scale_i * 2^(ref_i - lse2_i)
//     G_iy = (P_iy - 1) * dz_iy/dcos  (one entry per row, kept in fp32)
// so dx^ = diag(rho) (E' . w^) + G_iy w^_y   and   dw^ = E'^T . (diag(rho) x^) + sum_{i: y_i = j} G_iy x^_i.
// The two GEMMs run on the tensor cores (tc_head.cu); these kernels prepare rho / the scaled rows and add the sparse
// target-column terms (the closed-form backward of SURVEY.md section 8a, same maths as autograd of
// criterion.py:260-301 + nn.CrossEntropyLoss).
#include "common.cuh"

// one warp per row: rho_i, gty_i = G_iy, xs_i = bf16(rho_i * x^_i); rows in [B, B_pad) are zeroed.
__global__ void __launch_bounds__(256) stash_prep_kernel(const float* __restrict__ rowp, int64_t ldp,
                                                         const float* __restrict__ rowout, int64_t ldo,
                                                         const float* __restrict__ xhat32, int64_t B, int64_t B_pad,
                                                         float umax, __nv_bfloat16* __restrict__ xs,
                                                         float* __restrict__ rho, float* __restrict__ gty,
                                                         const int* __restrict__ fallback) {
  mh_pdl_sync();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= B_pad) return;
  uint2* dst = reinterpret_cast<uint2*>(xs + row * MH_D);
  if (row >= B) {
#pragma unroll
    for (int k = 0; k < 4; ++k) dst[lane + 32 * k] = make_uint2(0u, 0u);
    if (lane == 0) { rho[row] = 0.f; gty[row] = 0.f; }
    return;
  }
  // guarded stash, fallback taken (mh_step_backward): the B x C buffer holds the recomputed G itself, target column included
  const bool fb = fallback != nullptr && *fallback != 0;
  const float scale = rowp[MH_RP_SCALE * ldp + row];
  const float ref2 = scale * MH_LOG2E * umax - 102.f;
  const float r = fb ? 1.f : scale * exp2f(ref2 - rowout[MH_RO_LSE2 * ldo + row]);
  if (lane == 0) {
    rho[row] = r;
    gty[row] = fb ? 0.f : rowout[MH_RO_AUX0 * ldo + row] * rowp[MH_RP_DZT * ldp + row];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float4 v = reinterpret_cast<const float4*>(xhat32 + row * MH_D)[lane + 32 * k];
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x * r, v.y * r), p1 = __floats2bfloat162_rn(v.z * r, v.w * r);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&p0);
    pk.y = *reinterpret_cast<uint32_t*>(&p1);
    dst[lane + 32 * k] = pk;
  }
}

extern "C" int mh_stash_prep(const mh_config* cfg_host, const float* rowp, int64_t ldp, const float* rowout, int64_t ldo,
                             const float* x_hat32, int64_t B, int64_t B_pad, void* xs_bf16, float* rho, float* gty,
                             void* stream) {
  return mh_stash_prep_ex(cfg_host, rowp, ldp, rowout, ldo, x_hat32, B, B_pad, xs_bf16, rho, gty, nullptr, stream);
}

extern "C" int mh_stash_prep_ex(const mh_config* cfg_host, const float* rowp, int64_t ldp, const float* rowout, int64_t ldo,
                       const float* x_hat32, int64_t B, int64_t B_pad, void* xs_bf16, float* rho, float* gty,
                       const int* fallback, void* stream) {
  MH_CHECK_ARG(cfg_host && rowp && rowout && x_hat32 && xs_bf16 && rho && gty, "null pointer");
  MH_CHECK_ARG(B > 0 && B_pad >= B && ldp >= B && ldo >= B, "bad shape");
  const MhParams p = mh_make_params(cfg_host);
  mh_launch(stash_prep_kernel, (unsigned)((B_pad + 7) / 8), 256, 0, (cudaStream_t)stream, 
      rowp, ldp, rowout, ldo, x_hat32, B, B_pad, mh_family_umax(&p), (__nv_bfloat16*)xs_bf16, rho, gty, fallback);
  MH_LAUNCH_OK();
  return MH_OK;
}

// dx^_i = rho_i * sum_splits part_i + gty_i * w^_{y_i}  (the target term only on the shard that owns the label)
__global__ void __launch_bounds__(256) stash_dx_combine_kernel(const float* __restrict__ part, int n_split,
                                                               int64_t split_stride, const float* __restrict__ rho,
                                                               const float* __restrict__ gty,
                                                               const int32_t* __restrict__ label_local,
                                                               const __nv_bfloat16* __restrict__ what, int64_t B,
                                                               float* __restrict__ out) {
  mh_pdl_sync();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= B) return;
  // rho == NULL: plain sum of the split-K partials (recompute mode, before the cross-rank reduce-scatter)
  const float r = rho ? rho[row] : 1.f, g = rho ? gty[row] : 0.f;
  const int32_t y = rho ? label_local[row] : -1;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < n_split; ++s) {
      const float4 q = reinterpret_cast<const float4*>(part + (int64_t)s * split_stride + row * MH_D)[lane + 32 * k];
      a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
    }
    a.x *= r; a.y *= r; a.z *= r; a.w *= r;
    if (y >= 0) {
      const uint2 pk = reinterpret_cast<const uint2*>(what + (int64_t)y * MH_D)[lane + 32 * k];
      const float2 w0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.x));
      const float2 w1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.y));
      a.x = fmaf(g, w0.x, a.x); a.y = fmaf(g, w0.y, a.y); a.z = fmaf(g, w1.x, a.z); a.w = fmaf(g, w1.y, a.w);
    }
    reinterpret_cast<float4*>(out + row * MH_D)[lane + 32 * k] = a;
  }
}

extern "C" int mh_stash_dx_combine(const float* dxhat_part, int n_split, int64_t split_stride, const float* rho,
                                   const float* gty, const int32_t* label_local, const void* w_hat_bf16, int64_t B,
                                   float* dxhat, void* stream) {
  MH_CHECK_ARG(dxhat_part && dxhat, "null pointer");
  MH_CHECK_ARG(!rho || (gty && label_local && w_hat_bf16), "the stash terms need rho, gty, label_local and w_hat together");
  MH_CHECK_ARG(n_split >= 1 && B > 0, "bad shape");
  mh_launch(stash_dx_combine_kernel, (unsigned)((B + 7) / 8), 256, 0, (cudaStream_t)stream, 
      dxhat_part, n_split, split_stride, rho, gty, label_local, (const __nv_bfloat16*)w_hat_bf16, B, dxhat);
  MH_LAUNCH_OK();
  return MH_OK;
}

// dW_j += coef_j * (delta_j - w^_j (w^_j . delta_j)),  delta_j = sum_{i: y_i = j} gty_i x^_i,  coef_j = g / |w_j|.
// One warp per row; the first row of every label group gathers the whole group in row order and does a plain
// read-modify-write of dW_j (no atomics: bit-reproducible).
__global__ void __launch_bounds__(256) stash_dw_target_kernel(const float* __restrict__ gty,
                                                              const int32_t* __restrict__ label_local,
                                                              const float* __restrict__ xhat32,
                                                              const __nv_bfloat16* __restrict__ what,
                                                              const float* __restrict__ inv_norm,
                                                              const float* __restrict__ gscal, int64_t B, int layout,
                                                              float* __restrict__ dW, int64_t ld) {
  mh_pdl_sync();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= B) return;
  const int32_t y = label_local[row];
  if (y < 0) return;
  // label scans, 4 rows per lane and step (B = 8192 global rows on 8 GPUs: the one-row-per-lane scan took 0.18 ms);
  // bit e of the result: row i + e carries label y (rows >= limit masked off)
  auto scan4 = [&](int64_t i, int64_t limit) -> unsigned {
    unsigned h = 0;
    if (i + 3 < B) {
      const int4 v = __ldg(reinterpret_cast<const int4*>(label_local + i));
      h = (v.x == y ? 1u : 0u) | (v.y == y ? 2u : 0u) | (v.z == y ? 4u : 0u) | (v.w == y ? 8u : 0u);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (i + e < B && label_local[i + e] == y) h |= 1u << e;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (i + e >= limit) h &= ~(1u << e);
    return h;
  };
  // an earlier row with the same label owns the group
  for (int64_t i0 = 0; i0 < row; i0 += 128) {
    const unsigned h = scan4(i0 + 4 * lane, row);
    if (__ballot_sync(0xffffffffu, h != 0)) return;
  }
  float4 d[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) d[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  // gather the group in ascending row order (no earlier row matches, so starting at the 128-row block of `row` is exact)
  for (int64_t i0 = row & ~(int64_t)127; i0 < B; i0 += 128) {
    const unsigned h = scan4(i0 + 4 * lane, B);
    unsigned lm = __ballot_sync(0xffffffffu, h != 0);
    while (lm) {
      const int src = __ffs(lm) - 1;
      lm &= lm - 1;
      unsigned hh = __shfl_sync(0xffffffffu, h, src);
      while (hh) {
        const int e = __ffs(hh) - 1;
        hh &= hh - 1;
        const int64_t ii = i0 + 4 * src + e;
        const float g = gty[ii];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 v = reinterpret_cast<const float4*>(xhat32 + ii * MH_D)[lane + 32 * k];
          d[k].x = fmaf(g, v.x, d[k].x); d[k].y = fmaf(g, v.y, d[k].y);
          d[k].z = fmaf(g, v.z, d[k].z); d[k].w = fmaf(g, v.w, d[k].w);
        }
      }
    }
  }
  float4 w[4];
  float dot = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint2 pk = reinterpret_cast<const uint2*>(what + (int64_t)y * MH_D)[lane + 32 * k];
    const float2 w0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.x));
    const float2 w1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.y));
    w[k] = make_float4(w0.x, w0.y, w1.x, w1.y);
    dot += d[k].x * w[k].x + d[k].y * w[k].y + d[k].z * w[k].z + d[k].w * w[k].w;
  }
  dot = warp_sum(dot);
  const float coef = gscal[0] * inv_norm[y];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float4 o = make_float4((d[k].x - w[k].x * dot) * coef, (d[k].y - w[k].y * dot) * coef,
                                 (d[k].z - w[k].z * dot) * coef, (d[k].w - w[k].w * dot) * coef);
    const int dd = (lane + 32 * k) * 4;
    if (layout == MH_LAYOUT_CD) {
      float4* p = reinterpret_cast<float4*>(dW + (int64_t)y * ld + dd);
      float4 c = *p;
      c.x += o.x; c.y += o.y; c.z += o.z; c.w += o.w;
      *p = c;
    } else {
      dW[(int64_t)(dd + 0) * ld + y] += o.x;
      dW[(int64_t)(dd + 1) * ld + y] += o.y;
      dW[(int64_t)(dd + 2) * ld + y] += o.z;
      dW[(int64_t)(dd + 3) * ld + y] += o.w;
    }
  }
}

extern "C" int mh_stash_dw_target(const float* gty, const int32_t* label_local, const float* x_hat32,
                                  const void* w_hat_bf16, const float* inv_norm, const float* gscal, int64_t B,
                                  int layout, float* dW, int64_t ld, void* stream) {
  MH_CHECK_ARG(gty && label_local && x_hat32 && w_hat_bf16 && inv_norm && gscal && dW, "null pointer");
  MH_CHECK_ARG(B > 0, "bad shape");
  MH_CHECK_ARG(layout == MH_LAYOUT_CD || layout == MH_LAYOUT_DC, "unknown layout");
  MH_CHECK_ARG(layout != MH_LAYOUT_CD || (ld % 4 == 0 && ((uintptr_t)dW & 15) == 0), "CD dW must be 16-byte aligned");
  MH_CHECK_ARG(((uintptr_t)label_local & 15) == 0, "label_local must be 16-byte aligned");
  mh_launch(stash_dw_target_kernel, (unsigned)((B + 7) / 8), 256, 0, (cudaStream_t)stream, 
      gty, label_local, x_hat32, (const __nv_bfloat16*)w_hat_bf16, inv_norm, gscal, B, layout, dW, ld);
  MH_LAUNCH_OK();
  return MH_OK;
}
