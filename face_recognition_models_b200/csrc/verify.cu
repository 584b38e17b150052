// Pair verification: cosine of L2-normalised embedding pairs, the device-side piece of the reference's LFW evaluator
// (F.normalize(model(img1)) . F.normalize(model(img2)), main_code/utils/model_utils.py:333-335, 367-369, 392-394).
// HBM-bound: 2 * d * sizeof(T) bytes read per pair, 4 written.  One warp per pair, 16-byte loads when d % 8 == 0.
#include "common.cuh"

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

template <typename T>
__global__ void __launch_bounds__(256) pair_cosine_kernel(const T* __restrict__ e1, const T* __restrict__ e2, int64_t N,
                                                          int64_t d, int64_t ld1, int64_t ld2, float* __restrict__ out) {
  mh_pdl_sync();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= N) return;
  const T* a = e1 + row * ld1;
  const T* b = e2 + row * ld2;
  float aa = 0.f, bb = 0.f, ab = 0.f;
  constexpr int VEC = 16 / sizeof(T);
  const bool vec_ok = (d % VEC == 0) && (ld1 % VEC == 0) && (ld2 % VEC == 0) &&
                      ((reinterpret_cast<uintptr_t>(e1) | reinterpret_cast<uintptr_t>(e2)) & 15) == 0;
  if (vec_ok) {
    for (int64_t k = lane; k < d / VEC; k += 32) {
      const uint4 qa = __ldg(reinterpret_cast<const uint4*>(a) + k);
      const uint4 qb = __ldg(reinterpret_cast<const uint4*>(b) + k);
      const T* pa = reinterpret_cast<const T*>(&qa);
      const T* pb = reinterpret_cast<const T*>(&qb);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float x = to_f<T>(pa[e]), y = to_f<T>(pb[e]);
        aa = fmaf(x, x, aa); bb = fmaf(y, y, bb); ab = fmaf(x, y, ab);
      }
    }
  } else {
    for (int64_t k = lane; k < d; k += 32) {
      const float x = to_f<T>(a[k]), y = to_f<T>(b[k]);
      aa = fmaf(x, x, aa); bb = fmaf(y, y, bb); ab = fmaf(x, y, ab);
    }
  }
  aa = warp_sum(aa); bb = warp_sum(bb); ab = warp_sum(ab);
  // F.normalize: x / max(|x|, 1e-12) on both sides
  if (lane == 0) out[row] = ab / (fmaxf(sqrtf(aa), 1e-12f) * fmaxf(sqrtf(bb), 1e-12f));
}

extern "C" int mh_pair_cosine(const void* e1, const void* e2, int dtype, int64_t N, int64_t d, int64_t ld1, int64_t ld2,
                              float* cos_out, void* stream) {
  MH_CHECK_ARG(e1 && e2 && cos_out, "null pointer");
  MH_CHECK_ARG(N > 0 && d > 0 && ld1 >= d && ld2 >= d, "bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((N + 7) / 8));
  if (dtype == MH_F32)
    mh_launch(pair_cosine_kernel<float>, grid, 256, 0, st, (const float*)e1, (const float*)e2, N, d, ld1, ld2, cos_out);
  else if (dtype == MH_BF16)
    mh_launch(pair_cosine_kernel<__nv_bfloat16>, grid, 256, 0, st, (const __nv_bfloat16*)e1, (const __nv_bfloat16*)e2, N, d, ld1,
                                                            ld2, cos_out);
  else if (dtype == MH_F16)
    mh_launch(pair_cosine_kernel<__half>, grid, 256, 0, st, (const __half*)e1, (const __half*)e2, N, d, ld1, ld2, cos_out);
  else
    MH_CHECK_ARG(false, "unknown dtype");
  MH_LAUNCH_OK();
  return MH_OK;
}
