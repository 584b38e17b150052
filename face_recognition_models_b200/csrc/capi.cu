// C-ABI glue: error reporting, version, device check, hyper-parameter expansion.
#include "common.cuh"
#include <stdlib.h>
#include <cstdarg>
#include <cstdio>

static thread_local char g_err[512] = "";

void mh_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* mh_last_error(void) { return g_err; }
extern "C" const char* mh_version(void) { return "margin_head_b200 0.1.0 (sm_100a)"; }

extern "C" int mh_device_check(void) {
  int dev = 0;
  MH_CUDA_OK(cudaGetDevice(&dev));
  int major = 0;
  MH_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    mh_set_error("margin_head_b200 needs an sm_100 device (compute capability 10.x), found %d.x", major);
    return MH_ERR_UNSUPPORTED;
  }
  return MH_OK;
}

int mh_num_sms() {
  static MhDeviceOnce once;
  static int n_sm[MH_MAX_DEVICES];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MH_MAX_DEVICES) return 148;
  mh_once_per_device(once, [&] {
    int n = 0;
    cudaError_t e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    n_sm[dev] = (e == cudaSuccess && n > 0) ? n : 148;
    return cudaSuccess;
  });
  return n_sm[dev];
}

// Expand the reference constructor arguments into the constants every kernel needs.
// ArcFace criterion.py:246-249, CurricularFace :507-510, MV_Softmax :371-374; clamps per family.
MhParams mh_make_params(const mh_config* c) {
  MhParams p{};
  p.family = c->family;
  p.easy_margin = c->easy_margin;
  p.sphere_m = c->sphere_m;
  p.s = c->s;
  p.m = c->m;
  p.cos_m = (float)cos((double)c->m);
  p.sin_m = (float)sin((double)c->m);
  p.th = (float)cos(M_PI - (double)c->m);
  p.mm = (float)(sin(M_PI - (double)c->m) * (double)c->m);
  p.momentum = c->momentum;
  p.h = c->h;
  p.t_alpha = c->t_alpha;
  p.l_margin = c->l_margin;
  p.u_margin = c->u_margin;
  p.l_a = c->l_a;
  p.u_a = c->u_a;
  p.sphere_lambda = c->sphere_lambda;
  p.hard_kind = 0;
  p.hard_a = 0.f;
  p.hard_b = 0.f;
  p.scale_is_norm = 0;
  switch (c->family) {
    case MH_ARCFACE: p.lo = -INFINITY; p.hi = INFINITY; break;                       // no clamp, criterion.py:267-281
    case MH_COSFACE: p.lo = (float)(-1 + 1e-4); p.hi = (float)(1 - 1e-4); break;     // criterion.py:177
    case MH_SPHEREFACE: p.lo = -1.f; p.hi = 1.f; p.scale_is_norm = 1; break;          // criterion.py:81,105
    case MH_MV_AM:
    case MH_MV_ARC:
      p.lo = (float)(-1 + 1e-7); p.hi = (float)(1 - 1e-7);                            // criterion.py:413
      p.hard_kind = 1; p.hard_a = c->mv_weight; p.hard_b = c->mv_weight - 1.f;        // criterion.py:435
      break;
    case MH_CURRICULAR: p.lo = -1.f; p.hi = 1.f; p.hard_kind = 2; break;              // criterion.py:546,575
    case MH_ADAFACE: p.lo = (float)(-1 + 1e-3); p.hi = (float)(1 - 1e-3); break;      // criterion.py:872
    default: p.lo = (float)(-1 + 1e-7); p.hi = (float)(1 - 1e-7); break;              // criterion.py:994,1104,1260
  }
  return p;
}

bool mh_pdl_enabled() {
  // off by default: measured on B200 it costs BASELINE config 3 0.53 -> 0.61 ms per step and changes nothing at configs 2
  // and 4 (profiles/r2_ab_pdl_flush.txt); MH_PDL=1 turns it on
  static const bool on = [] { const char* e = getenv("MH_PDL"); return e && e[0] == '1'; }();
  return on;
}

float mh_family_umax(const MhParams* p) {
  if (p->hard_kind == 1) return fmaxf(1.f, p->hard_a + p->hard_b);   // MV: w*c + w - 1 at c = 1 (criterion.py:435)
  if (p->hard_kind == 2) return 2.f;                                  // Curricular: c*(t+c), t <= 1 (criterion.py:575)
  return 1.f;
}
