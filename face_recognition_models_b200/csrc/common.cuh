// Shared helpers for the margin-head kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <math.h>
#include <mutex>
#include <utility>
#include "../../include/margin_head.h"

#define MH_LOG2E 1.4426950408889634f
#define MH_LN2 0.6931471805599453f

void mh_set_error(const char* fmt, ...);

#define MH_CHECK_ARG(cond, msg)                              \
  do {                                                       \
    if (!(cond)) {                                           \
      mh_set_error("%s: %s", __func__, msg);                 \
      return MH_ERR_ARG;                                     \
    }                                                        \
  } while (0)

#define MH_CUDA_OK(expr)                                                            \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      mh_set_error("%s: %s -> %s", __func__, #expr, cudaGetErrorString(_e));        \
      return MH_ERR_CUDA;                                                           \
    }                                                                               \
  } while (0)

#define MH_LAUNCH_OK()                                                              \
  do {                                                                              \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      mh_set_error("%s: launch failed: %s", __func__, cudaGetErrorString(_e));      \
      return MH_ERR_CUDA;                                                           \
    }                                                                               \
  } while (0)

// Per-device one-time host state (kernel attributes are per device; so is the SM count): one once-flag per device
// ordinal, so the entry points stay re-entrant and a process may drive several GPUs ("no globals
// except immutable kernel attributes set once under std::call_once").
constexpr int MH_MAX_DEVICES = 64;
struct MhDeviceOnce {
  std::once_flag flag[MH_MAX_DEVICES];
  cudaError_t err[MH_MAX_DEVICES];
};
template <class Fn>
inline cudaError_t mh_once_per_device(MhDeviceOnce& o, Fn&& fn) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= MH_MAX_DEVICES) return fn();          // beyond the cache: set it every time (cheap)
  std::call_once(o.flag[dev], [&] { o.err[dev] = fn(); });
  return o.err[dev];
}
// SM count of the CURRENT device (cached per device ordinal; 148 if the query fails).
int mh_num_sms();

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------
// A step is a chain of ~15 dependent launches, most of them O(B*d) kernels of a few microseconds; BASELINE configs 2 and 3
// are bounded by that chain.  Every kernel of this library is launched with the programmatic-stream-serialization
// attribute and starts with mh_pdl_sync(): `griddepcontrol.wait` (all prerequisite grids complete, their writes visible)
// followed by `griddepcontrol.launch_dependents` (the next kernel of the stream may be scheduled now; it parks in its own
// wait).  Launch latency and CTA ramp-up of kernel N+1 thus overlap the body of kernel N, while the data dependencies
// stay exactly those of plain stream order: nothing is read or written before the wait, and a kernel only triggers
// after its own wait, so completion is transitive along the chain.  Kernels of other libraries (torch, NCCL) in between
// never trigger early, so the attribute degrades to ordinary serialization next to them.  OPT-IN (MH_PDL=1): measured,
// it does not pay on this part -- config 3 gets slower, configs 2 and 4 do not move -- so by
// default no launch carries the attribute and the two instructions at the top of every kernel are no-ops.
bool mh_pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void mh_pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... Params, typename... Args>
inline cudaError_t mh_launch(void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = mh_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<Params>(std::forward<Args>(args))...);
}
#endif

// Device copy of the hyper-parameters + derived constants, passed by value to kernels.
struct MhParams {
  int family;
  int easy_margin;
  int sphere_m;
  int hard_kind;          // 0 none, 1 MV (c>thr -> a*c+b), 2 Curricular (c>thr -> c*(t_buf+c), t_buf read from state)
  float s, m;
  float cos_m, sin_m, th, mm;
  float lo, hi;           // clamp bounds on the raw cosine (ArcFace: -inf/+inf)
  float hard_a, hard_b;
  float momentum, h, t_alpha;
  float l_margin, u_margin, l_a, u_a;
  float sphere_lambda;
  int scale_is_norm;      // SphereFace: logit scale is |x_i|
};

MhParams mh_make_params(const mh_config* c);
// Upper bound of u = z / scale over non-target columns, per family (fixed-reference softmax, see mh_tc_fixref_ok).
float mh_family_umax(const MhParams* p);

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Element transform shared by every B x C kernel (tensor-core and exact paths).
// raw cosine -> (u, du/dcos_raw, c_clamped); the row scale multiplies outside.
// MV-Softmax criterion.py:433-435, CurricularFace criterion.py:559,575, clamps per family.
struct ElemOut { float u, du, c; };
__device__ __forceinline__ ElemOut mh_elem(float raw, float lo, float hi, int hard_kind, float thr,
                                           float ha, float hb) {
  ElemOut o;
  float c = fminf(fmaxf(raw, lo), hi);
  float inside = (raw >= lo && raw <= hi) ? 1.f : 0.f;
  o.c = c;
  if (hard_kind == 1 && c > thr) {
    o.u = fmaf(ha, c, hb);
    o.du = ha * inside;
  } else if (hard_kind == 2 && c > thr) {
    o.u = c * (ha + c);
    o.du = (ha + 2.f * c) * inside;
  } else {
    o.u = c;
    o.du = inside;
  }
  return o;
}
