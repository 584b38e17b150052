// One C call per phase of a training step (single GPU, tensor-core path): mh_step_forward / mh_step_backward enqueue the
// whole kernel sequence that face_recognition_models_b200/functional.py otherwise drives entry point by entry point
// (~20 FFI calls per step).  BASELINE configs 2 and 3 (B=512, C=10,575 / B=1024, C=85,742) are launch-bound: their
// kernels sum to 0.1-0.4 ms while the per-call host overhead of the Python driver was 0.4-0.6 ms.  Nothing new runs on the
// device here: these are the same entry points of include/margin_head.h in the same order, sharing one workspace
// descriptor (mh_step_ws, all caller-owned device pointers).
#include "common.cuh"

#define STEP_TRY(call)                 \
  do {                                 \
    if (int _e = (call)) return _e;    \
  } while (0)

static int check_ws(const mh_step_ws* ws) {
  MH_CHECK_ARG(ws, "null workspace descriptor");
  MH_CHECK_ARG(ws->B > 0 && ws->B_pad >= ws->B && ws->B_pad % 256 == 0, "B_pad must be a multiple of 256 and >= B");
  MH_CHECK_ARG(ws->C > 0 && ws->C_pad >= ws->C && ws->C_pad % 256 == 0, "C_pad must be a multiple of 256 and >= C");
  MH_CHECK_ARG(ws->w_hat && ws->inv_norm && ws->x_hat && ws->x_hat32 && ws->xnorm && ws->t_raw && ws->label_local &&
                   ws->rowp && ws->stats_tiles && ws->merge_scratch && ws->stats && ws->rowout,
               "null forward workspace pointer");
  return MH_OK;
}

static int step_forward_body(const mh_config* cfg, const mh_step_ws* ws, const void* x, const int64_t* labels,
                             const float* W, const float* margins, float* state, int update_state, int run_prologue_w,
                             int stash, float* scalars, void* stream) {
  MH_CHECK_ARG(cfg && x && labels && W && state && scalars, "null pointer");
  STEP_TRY(check_ws(ws));
  MH_CHECK_ARG(!stash || ws->bc, "stash requested without a B x C buffer");
  MH_CHECK_ARG(stash >= 0 && stash <= 2, "stash must be 0 (none), 1 (proven) or 2 (guarded)");
  MH_CHECK_ARG(stash != 2 || ws->guard, "the guarded stash needs ws->guard");
  MH_CHECK_ARG(stash != 2 || !ws->pw_ready, "the guarded stash does not combine with the merged prologue + forward");
  MH_CHECK_ARG(ws->n_tiles == mh_fwd_num_tiles(ws->C_pad), "stats_tiles must hold mh_fwd_num_tiles records");
  // merged prologue + forward (ws->pw_ready set, the W prologue has to run, eligible head / shape): the prologue becomes
  // a role of the forward launch; 1/|w_y| of the target rows is then taken by the x prologue from the gathered rows
  int pw_ok = 0;
  if (run_prologue_w && ws->pw_ready)
    STEP_TRY(mh_tc_forward_pw(cfg, nullptr, ws->B, ws->B_pad, W, ws->layout, ws->ld, nullptr, nullptr, ws->C, ws->C_pad, nullptr,
                              ws->B_pad, nullptr, nullptr, nullptr, stash ? ws->bc : nullptr, nullptr, &pw_ok, stream));
  if (run_prologue_w && !pw_ok)
    STEP_TRY(mh_prologue_w(W, ws->layout, ws->C, ws->ld, ws->w_hat, ws->C_pad, nullptr, ws->inv_norm, stream));
  STEP_TRY(mh_prologue_x(x, ws->x_dtype, ws->B, ws->B_pad, labels, W, ws->layout, ws->C, ws->ld, /*c_offset=*/0,
                         pw_ok ? nullptr : ws->inv_norm, ws->x_hat, ws->x_hat32, ws->xnorm, ws->t_raw, ws->label_local,
                         /*c_total=*/ws->C, stream));
  STEP_TRY(mh_row_params(cfg, ws->B, ws->xnorm, ws->t_raw, margins, state, update_state, ws->rowp, ws->B_pad, stream));
  if (pw_ok)
    STEP_TRY(mh_tc_forward_pw(cfg, ws->x_hat, ws->B, ws->B_pad, W, ws->layout, ws->ld, ws->w_hat, ws->inv_norm, ws->C, ws->C_pad,
                              ws->rowp, ws->B_pad, ws->label_local, state, ws->stats_tiles, stash ? ws->bc : nullptr,
                              ws->pw_ready, &pw_ok, stream));
  else
    STEP_TRY(mh_tc_forward_ex(cfg, ws->x_hat, ws->B, ws->B_pad, ws->w_hat, ws->C, ws->C_pad, ws->rowp, ws->B_pad,
                                ws->label_local, state, ws->stats_tiles, stash ? ws->bc : nullptr, stash, nullptr, 0, stream));
  const int sphere = cfg->family == MH_SPHEREFACE ? 1 : 0;
  STEP_TRY(mh_merge_stats(ws->stats_tiles, ws->n_tiles, ws->B, ws->B_pad, ws->merge_scratch, ws->stats, stream));
  if (stash != 2) {
    STEP_TRY(mh_finalize_rows(ws->stats, ws->B_pad, ws->rowp, ws->B_pad, ws->B, ws->B, sphere, ws->rowout, ws->B_pad, scalars,
                              state, stream));
    return MH_OK;
  }
  // Guarded stash: the finaliser decides on the device whether the speculative fixed-reference sums stand (every row sum
  // >= C 2^-102, see mh_tc_stash_guarded_ok) and sets ws->guard; the general forward follows as launches gated on it.
  const float guard_min_l = ldexpf((float)ws->C, -102);
  STEP_TRY(mh_finalize_rows_ex(ws->stats, ws->B_pad, ws->rowp, ws->B_pad, ws->B, ws->B, sphere, ws->rowout, ws->B_pad,
                                 scalars, state, guard_min_l, ws->guard, nullptr, 0, stream));
  STEP_TRY(mh_tc_forward_ex(cfg, ws->x_hat, ws->B, ws->B_pad, ws->w_hat, ws->C, ws->C_pad, ws->rowp, ws->B_pad,
                              ws->label_local, state, ws->stats_tiles, nullptr, 0, ws->guard, 1, stream));
  STEP_TRY(mh_merge_stats_ex(ws->stats_tiles, ws->n_tiles, ws->B, ws->B_pad, ws->merge_scratch, ws->stats, ws->guard, 1,
                               stream));
  STEP_TRY(mh_finalize_rows_ex(ws->stats, ws->B_pad, ws->rowp, ws->B_pad, ws->B, ws->B, sphere, ws->rowout, ws->B_pad,
                                 scalars, state, 0.f, nullptr, ws->guard, 1, stream));
  return MH_OK;
}

static int step_backward_body(const mh_config* cfg, const mh_step_ws* ws, int stash, const float* state,
                              const float* g_loss, const float* g_lossg, void* dx, float* dW, void* stream) {
  MH_CHECK_ARG(cfg && state, "null pointer");
  STEP_TRY(check_ws(ws));
  MH_CHECK_ARG(ws->bc && ws->gscal && ws->dxhat_part, "null backward workspace pointer");
  MH_CHECK_ARG(ws->r_colsum || (ws->rpart && ws->rflag), "need r_colsum (side-pass projection) or rpart + rflag (self-projection)");
  MH_CHECK_ARG(!stash || (ws->xs && ws->rho && ws->gty && ws->dxhat_full), "null stash workspace pointer");
  MH_CHECK_ARG(stash != 2 || (ws->guard && !ws->r_colsum && ws->rpart && ws->rflag),
               "the guarded stash needs ws->guard and the self-projecting dW kernel (r_colsum == NULL, rpart, rflag)");
  const int* fallback = stash == 2 ? ws->guard : nullptr;
  const float* rowout = ws->rowout;
  const float* aux0 = rowout + (int64_t)MH_RO_AUX0 * ws->B_pad;
  const float* aux1 = rowout + (int64_t)MH_RO_AUX1 * ws->B_pad;
  const float* lse2 = rowout + (int64_t)MH_RO_LSE2 * ws->B_pad;
  // projection term r_j = w^_j . dw^_j of the dW epilogue: either produced beforehand (stash: the dx kernel's side pass,
  // one partial per 128-row block; recompute: the backward-G column sums) or taken from the dW accumulators themselves
  const bool selfp = ws->r_colsum == nullptr;
  const int r_parts = stash ? (int)(ws->B_pad / MH_TILE) : 1;
  STEP_TRY(mh_make_gscal(g_loss, g_lossg, ws->B, ws->gscal, stream));
  int n_split = 0;
  STEP_TRY(mh_tc_backward_dx(nullptr, ws->B_pad, ws->C_pad, nullptr, nullptr, &n_split, nullptr, stream));
  MH_CHECK_ARG(n_split <= ws->part_splits, "dxhat_part holds fewer splits than mh_tc_backward_dx needs");
  const int64_t split_stride = ws->B_pad * MH_D;
  const void* xs = ws->x_hat;
  // merged dx + dW kernel (ws->prog set, both gradients wanted, eligible shape): one launch for both backward GEMMs
  int merged_split = 0;
  if (ws->prog && dx && dW && ws->rpart && ws->rflag)
    STEP_TRY(mh_tc_backward_dxdw(nullptr, ws->B_pad, ws->C, ws->C_pad, nullptr, nullptr, nullptr, nullptr, ws->layout, nullptr,
                                 ws->ld, nullptr, &merged_split, nullptr, nullptr, nullptr, stream));
  if (merged_split > 0) {
    MH_CHECK_ARG(merged_split <= ws->part_splits, "dxhat_part holds fewer splits than the merged backward needs");
    if (stash) {
      if (fallback)      // guarded stash whose forward fell back: rewrite the stash with the recomputed G (no-op otherwise)
        STEP_TRY(mh_tc_backward_g_ex(cfg, ws->x_hat, ws->B, ws->B_pad, ws->w_hat, ws->C, ws->C_pad, ws->rowp, ws->B_pad,
                                       ws->label_local, state, lse2, ws->bc, nullptr, fallback, 1, stream));
      STEP_TRY(mh_stash_prep_ex(cfg, ws->rowp, ws->B_pad, rowout, ws->B_pad, ws->x_hat32, ws->B, ws->B_pad, ws->xs, ws->rho,
                                  ws->gty, fallback, stream));
      xs = ws->xs;
    } else {
      STEP_TRY(mh_tc_backward_g(cfg, ws->x_hat, ws->B, ws->B_pad, ws->w_hat, ws->C, ws->C_pad, ws->rowp, ws->B_pad,
                                ws->label_local, state, lse2, ws->bc, nullptr, stream));
    }
    STEP_TRY(mh_tc_backward_dxdw(ws->bc, ws->B_pad, ws->C, ws->C_pad, ws->w_hat, xs, ws->inv_norm, ws->gscal, ws->layout, dW,
                                 ws->ld, ws->dxhat_part, &merged_split, ws->rpart, ws->rflag, ws->prog, stream));
    if (stash) {
      STEP_TRY(mh_stash_dx_combine(ws->dxhat_part, merged_split, split_stride, ws->rho, ws->gty, ws->label_local, ws->w_hat,
                                   ws->B, ws->dxhat_full, stream));
      STEP_TRY(mh_norm_backward_x(ws->dxhat_full, 1, split_stride, ws->x_hat32, ws->xnorm, ws->rowp, ws->B_pad, aux0, aux1,
                                  ws->gscal, ws->B, dx, ws->x_dtype, stream));
      STEP_TRY(mh_stash_dw_target(ws->gty, ws->label_local, ws->x_hat32, ws->w_hat, ws->inv_norm, ws->gscal, ws->B,
                                  ws->layout, dW, ws->ld, stream));
    } else {
      STEP_TRY(mh_norm_backward_x(ws->dxhat_part, merged_split, split_stride, ws->x_hat32, ws->xnorm, ws->rowp, ws->B_pad,
                                  aux0, aux1, ws->gscal, ws->B, dx, ws->x_dtype, stream));
    }
    return MH_OK;
  }
  if (stash) {
    if (fallback)
      STEP_TRY(mh_tc_backward_g_ex(cfg, ws->x_hat, ws->B, ws->B_pad, ws->w_hat, ws->C, ws->C_pad, ws->rowp, ws->B_pad,
                                     ws->label_local, state, lse2, ws->bc, nullptr, fallback, 1, stream));
    STEP_TRY(mh_stash_prep_ex(cfg, ws->rowp, ws->B_pad, rowout, ws->B_pad, ws->x_hat32, ws->B, ws->B_pad, ws->xs, ws->rho,
                                ws->gty, fallback, stream));
    xs = ws->xs;
    if (selfp) {
      if (dx) STEP_TRY(mh_tc_backward_dx(ws->bc, ws->B_pad, ws->C_pad, ws->w_hat, ws->dxhat_part, &n_split, ws->dx_sync, stream));
    } else if (dx || dW) {          // the side pass of this GEMM feeds dW's projection: it runs even when only dW is wanted
      STEP_TRY(mh_tc_backward_dx_stash(cfg, ws->bc, ws->B_pad, ws->C, ws->C_pad, ws->w_hat, ws->rho, ws->rowp, ws->B_pad,
                                       ws->dxhat_part, ws->r_colsum, &n_split, ws->dx_sync, stream));
    }
    if (dx) {
      STEP_TRY(mh_stash_dx_combine(ws->dxhat_part, n_split, split_stride, ws->rho, ws->gty, ws->label_local, ws->w_hat,
                                   ws->B, ws->dxhat_full, stream));
      STEP_TRY(mh_norm_backward_x(ws->dxhat_full, 1, split_stride, ws->x_hat32, ws->xnorm, ws->rowp, ws->B_pad, aux0, aux1,
                                  ws->gscal, ws->B, dx, ws->x_dtype, stream));
    }
  } else {
    STEP_TRY(mh_tc_backward_g(cfg, ws->x_hat, ws->B, ws->B_pad, ws->w_hat, ws->C, ws->C_pad, ws->rowp, ws->B_pad,
                              ws->label_local, state, lse2, ws->bc, (dW && !selfp) ? ws->r_colsum : nullptr, stream));
    if (dx) {
      STEP_TRY(mh_tc_backward_dx(ws->bc, ws->B_pad, ws->C_pad, ws->w_hat, ws->dxhat_part, &n_split, ws->dx_sync, stream));
      STEP_TRY(mh_norm_backward_x(ws->dxhat_part, n_split, split_stride, ws->x_hat32, ws->xnorm, ws->rowp, ws->B_pad, aux0,
                                  aux1, ws->gscal, ws->B, dx, ws->x_dtype, stream));
    }
  }
  if (dW) {
    if (selfp)
      STEP_TRY(mh_tc_backward_dw_proj(ws->bc, ws->B_pad, ws->C, ws->C_pad, xs, ws->w_hat, ws->inv_norm, ws->gscal,
                                      ws->layout, dW, ws->ld, ws->rpart, ws->rflag, stream));
    else
      STEP_TRY(mh_tc_backward_dw_fused(ws->bc, ws->B_pad, ws->C, ws->C_pad, xs, ws->w_hat, ws->inv_norm, ws->r_colsum,
                                       r_parts, ws->gscal, ws->layout, dW, ws->ld, stream));
    if (stash)
      STEP_TRY(mh_stash_dw_target(ws->gty, ws->label_local, ws->x_hat32, ws->w_hat, ws->inv_norm, ws->gscal, ws->B,
                                  ws->layout, dW, ws->ld, stream));
  }
  return MH_OK;
}


// ------------------------------------------------------------------------------------------------------------------
// Graph replay of a phase.  A phase is 7-12 dependent launches, most of them a few microseconds long; at BASELINE
// configs 2 and 3 the gaps between them and the host time to issue them are a large part of the step.  When the caller
// hands in a cache (ws->graph_cache, mh_step_cache_create), a phase is captured the first time a set of arguments
// (hyper-parameters, workspace descriptor, every pointer and flag, compared byte by byte) is seen -- on the cache's
// private stream, so the caller's stream may be the legacy default stream -- and replayed with ONE cudaGraphLaunch on
// the caller's stream whenever that set comes back: same kernels, same arguments, same order, hence the same bits.  Any change of
// an argument (a fresh output tensor at another address, SphereFace's annealed lambda) is just a different key; the
// cache holds a few keys (LRU), and a phase whose key keeps changing falls back to plain launches for a while instead
// of re-capturing every step.  A caller that is itself capturing (tests/test_gpu_graph.py) gets plain launches, which
// its own capture records.  MH_STEP_GRAPH=0 disables the mechanism.
// ------------------------------------------------------------------------------------------------------------------
#include <string.h>
#include <stdlib.h>
#include <vector>

namespace {

struct PhaseKey {
  mh_config cfg;
  mh_step_ws ws;
  const void* p[8];
  int i[4];
};

struct PhaseEntry {
  PhaseKey key;
  cudaGraphExec_t exec;
  uint64_t last_use;
};

struct PhaseCache {
  std::vector<PhaseEntry> entries;
  int misses_in_row = 0;
  int64_t cool_down = 0;          // calls left during which this phase does not try to capture
};

constexpr int kMaxEntries = 6;
constexpr int kMissLimit = 6;      // captures in a row without a hit before the phase backs off
constexpr int kCoolDown = 200;

}  // namespace

struct mh_step_cache_s {
  std::mutex mu;
  int device = -1;
  cudaStream_t cap_stream = nullptr;
  uint64_t tick = 0;
  int64_t n_replay = 0, n_capture = 0, n_plain = 0;
  PhaseCache phase[2];             // 0 forward, 1 backward
};

static bool step_graphs_enabled() {
  static const bool on = [] { const char* e = getenv("MH_STEP_GRAPH"); return !(e && e[0] == '0'); }();
  return on;
}

extern "C" int mh_step_cache_create(void** out) {
  MH_CHECK_ARG(out, "null pointer");
  *out = nullptr;
  int dev = -1;
  MH_CUDA_OK(cudaGetDevice(&dev));
  cudaStream_t cap = nullptr;
  MH_CUDA_OK(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
  mh_step_cache_s* c = new mh_step_cache_s();
  c->device = dev;
  c->cap_stream = cap;
  *out = c;
  return MH_OK;
}

extern "C" int mh_step_cache_destroy(void* cache) {
  if (!cache) return MH_OK;
  mh_step_cache_s* c = static_cast<mh_step_cache_s*>(cache);
  {
    std::lock_guard<std::mutex> lk(c->mu);
    for (int ph = 0; ph < 2; ++ph) {
      for (PhaseEntry& en : c->phase[ph].entries) cudaGraphExecDestroy(en.exec);
      c->phase[ph].entries.clear();
    }
    if (c->cap_stream) cudaStreamDestroy(c->cap_stream);
    c->cap_stream = nullptr;
  }
  delete c;
  return MH_OK;
}

extern "C" int mh_step_cache_stats(void* cache, int64_t* counts3) {
  MH_CHECK_ARG(cache && counts3, "null pointer");
  mh_step_cache_s* c = static_cast<mh_step_cache_s*>(cache);
  std::lock_guard<std::mutex> lk(c->mu);
  counts3[0] = c->n_replay;
  counts3[1] = c->n_capture;
  counts3[2] = c->n_plain;
  return MH_OK;
}

// Runs `body(stream)` either directly or through the cache.  Returns the body's status (a failed capture falls back to
// direct launches; only the direct run's status is reported).
template <class Body>
static int run_phase(void* cache_v, int ph, const PhaseKey& key, void* stream, Body&& body) {
  mh_step_cache_s* c = static_cast<mh_step_cache_s*>(cache_v);
  if (!c || !step_graphs_enabled()) return body(stream);
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev != c->device) return body(stream);
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing((cudaStream_t)stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return body(stream);                                   // the caller is capturing: let its capture record the launches
  }
  std::lock_guard<std::mutex> lk(c->mu);
  PhaseCache& pc = c->phase[ph];
  ++c->tick;
  for (PhaseEntry& en : pc.entries) {
    if (memcmp(&en.key, &key, sizeof(PhaseKey)) == 0) {
      en.last_use = c->tick;
      pc.misses_in_row = 0;
      ++c->n_replay;
      MH_CUDA_OK(cudaGraphLaunch(en.exec, (cudaStream_t)stream));
      return MH_OK;
    }
  }
  // unknown key: capture it now, unless this phase keeps producing new keys (fresh addresses or hyper-parameters on
  // every call): after kMissLimit captures in a row without a single hit it issues plain launches for kCoolDown calls
  if (pc.cool_down > 0) {
    --pc.cool_down;
    ++c->n_plain;
    return body(stream);
  }
  if (pc.misses_in_row >= kMissLimit) {
    pc.misses_in_row = 0;
    pc.cool_down = kCoolDown;
    ++c->n_plain;
    return body(stream);
  }
  ++pc.misses_in_row;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  bool ok = cudaStreamBeginCapture(c->cap_stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
  if (ok) {
    const int rc = body((void*)c->cap_stream);
    const cudaError_t ee = cudaStreamEndCapture(c->cap_stream, &graph);
    ok = rc == MH_OK && ee == cudaSuccess && graph != nullptr;
  }
  if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
  if (graph) cudaGraphDestroy(graph);
  if (!ok) {
    cudaGetLastError();
    pc.cool_down = kCoolDown;                               // something in this phase does not capture: plain launches
    ++c->n_plain;
    return body(stream);
  }
  if ((int)pc.entries.size() >= kMaxEntries) {
    size_t lru = 0;
    for (size_t k = 1; k < pc.entries.size(); ++k)
      if (pc.entries[k].last_use < pc.entries[lru].last_use) lru = k;
    cudaGraphExecDestroy(pc.entries[lru].exec);
    pc.entries.erase(pc.entries.begin() + (long)lru);
  }
  PhaseEntry en;
  en.key = key;
  en.exec = exec;
  en.last_use = c->tick;
  pc.entries.push_back(en);
  ++c->n_capture;
  MH_CUDA_OK(cudaGraphLaunch(exec, (cudaStream_t)stream));
  return MH_OK;
}

extern "C" int mh_step_forward(const mh_config* cfg, const mh_step_ws* ws, const void* x, const int64_t* labels,
                               const float* W, const float* margins, float* state, int update_state, int run_prologue_w,
                               int stash, float* scalars, void* stream) {
  if (!cfg || !ws || !ws->graph_cache)
    return step_forward_body(cfg, ws, x, labels, W, margins, state, update_state, run_prologue_w, stash, scalars, stream);
  PhaseKey key;
  memset(&key, 0, sizeof(key));
  key.cfg = *cfg;
  key.ws = *ws;
  key.p[0] = x; key.p[1] = labels; key.p[2] = W; key.p[3] = margins; key.p[4] = state; key.p[5] = scalars;
  key.i[0] = update_state; key.i[1] = run_prologue_w; key.i[2] = stash;
  return run_phase(ws->graph_cache, 0, key, stream, [&](void* st) {
    return step_forward_body(cfg, ws, x, labels, W, margins, state, update_state, run_prologue_w, stash, scalars, st);
  });
}

extern "C" int mh_step_backward(const mh_config* cfg, const mh_step_ws* ws, int stash, const float* state,
                                const float* g_loss, const float* g_lossg, void* dx, float* dW, void* stream) {
  if (!cfg || !ws || !ws->graph_cache) return step_backward_body(cfg, ws, stash, state, g_loss, g_lossg, dx, dW, stream);
  PhaseKey key;
  memset(&key, 0, sizeof(key));
  key.cfg = *cfg;
  key.ws = *ws;
  key.p[0] = state; key.p[1] = g_loss; key.p[2] = g_lossg; key.p[3] = dx; key.p[4] = dW;
  key.i[0] = stash;
  return run_phase(ws->graph_cache, 1, key, stream, [&](void* st) {
    return step_backward_body(cfg, ws, stash, state, g_loss, g_lossg, dx, dW, st);
  });
}
