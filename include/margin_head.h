/*
 * margin_head.h - C ABI of the B200-native large-margin cosine-softmax head.
 *
 * The reference (Lac-quan-yeu-doi/Face-Recognition-Models) has no FFI: the hot path is a set of
 * Python nn.Modules in main_code/utils/criterion.py called from main_code/utils/model_utils.py:177.
 * This header is the boundary a maintainer binds with ctypes: plain pointers
 * and sizes, no torch types.  Every pointer is a DEVICE pointer unless its name ends in _host.
 * Every function enqueues work on `stream` (a cudaStream_t passed as void*), never allocates,
 * never synchronises, and returns 0 on success or a negative mh_status; mh_last_error() gives a
 * thread-local message.  All kernels are compiled for sm_100a only; there is no CPU fallback.
 *
 * Path:   x[B,512], W -> normalise -> cos = x^ w^T -> clamp -> margin on the
 * target column / hard-negative re-weighting -> scale -> softmax cross-entropy (+ top-1/5 rank)
 * -> dx, dW.
 */
#ifndef MARGIN_HEAD_H_
#define MARGIN_HEAD_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MH_D 512                 /* embedding dimension (config.py:13 FEATURE_DIM)        */
#define MH_TILE 128              /* B and C are padded to multiples of this in workspaces */
#define MH_NTILE_FWD 256         /* class-tile width of the tensor-core forward           */

typedef enum mh_status {
  MH_OK = 0,
  MH_ERR_ARG = -1,               /* bad shape / null pointer / misalignment */
  MH_ERR_CUDA = -2,              /* CUDA runtime or driver error            */
  MH_ERR_UNSUPPORTED = -3        /* e.g. device is not sm_100               */
} mh_status;

/* One value per reference head class (criterion.py line of the class in the comment). */
typedef enum mh_family {
  MH_ARCFACE = 0,                /* criterion.py:232  */
  MH_COSFACE = 1,                /* criterion.py:137  */
  MH_SPHEREFACE = 2,             /* criterion.py:12   */
  MH_MV_AM = 3,                  /* criterion.py:327, margin_type='am'  */
  MH_MV_ARC = 4,                 /* criterion.py:327, margin_type='arc' */
  MH_CURRICULAR = 5,             /* criterion.py:491  */
  MH_ADAFACE = 6,                /* criterion.py:795  */
  MH_ELASTIC_COS = 7,            /* criterion.py:951  */
  MH_ELASTIC_ARC = 8,            /* criterion.py:1054 */
  MH_MAGFACE = 9,                /* criterion.py:1178 */
  MH_VPL_ARC = 10                /* criterion.py:619 (VPLArcFace: memory-bank virtual proxies) */
} mh_family;

typedef enum mh_layout {
  MH_LAYOUT_CD = 0,              /* parameter `weight [C, D]` (ArcFace, SphereFace, MV_Softmax) */
  MH_LAYOUT_DC = 1               /* parameter `kernel [D, C]` (all other heads)                */
} mh_layout;

typedef enum mh_dtype { MH_F32 = 0, MH_BF16 = 1, MH_F16 = 2 } mh_dtype;

/* Row-parameter planes produced by mh_row_params: rowp[plane * ldp + row], float32. */
enum {
  MH_RP_SCALE = 0,    /* logit scale of the row: s, or |x_i| for SphereFace (criterion.py:105)        */
  MH_RP_THR = 1,      /* hard-negative threshold (MV: criterion.py:424/430, Curricular: :559), +inf   */
  MH_RP_ZT = 2,       /* target logit z_{i,y_i}                                                       */
  MH_RP_DZT = 3,      /* d z_{i,y_i} / d cos_raw_{i,y_i} (clamp mask folded in)                       */
  MH_RP_T = 4,        /* clamped target cosine (pre-margin value at the target column)                */
  MH_RP_DZT_DN = 5,   /* d z_{i,y_i} / d |x_i| through the margin (MagFace, criterion.py:1264)        */
  MH_RP_DLG_DN = 6,   /* d loss_g / d |x_i| (MagFace, criterion.py:1235-1238)                         */
  MH_RP_NORMS = 7,    /* what the reference returns as `norms` (MagFace: clamped, criterion.py:1290)  */
  MH_RP_PLANES = 8
};

/* Per-row forward statistics, stats[plane * lds + row], float32 (log2 domain for M/L). */
enum {
  MH_ST_M = 0,        /* running max of z*log2(e)                                      */
  MH_ST_L = 1,        /* sum of exp2(z*log2(e) - M)                                    */
  MH_ST_CNT = 2,      /* #non-target classes with pre-margin logit > target's (metrics.py:8) */
  MH_ST_EZ = 3,       /* sum of exp2(z*log2e - M) * u   (SphereFace d|x| term), else 0 */
  MH_ST_PLANES = 4
};

/* Per-row results of mh_combine, rowout[plane * ldo + row], float32. */
enum {
  MH_RO_LSE2 = 0,     /* log2-sum-exp2 of the row's logits (global over all shards)        */
  MH_RO_LOSS = 1,     /* lse - z_target (natural log units)                                 */
  MH_RO_CNT = 2,      /* rank count                                                        */
  MH_RO_AUX0 = 3,     /* d loss_row / d z_target = P_iy - 1 (feeds the MagFace |x| path)   */
  MH_RO_AUX1 = 4,     /* sum_j (P_ij - Y_ij) u_ij (SphereFace |x| path), else 0            */
  MH_RO_PLANES = 5
};

/* Hyper-parameters: the reference constructor arguments (criterion.py:17-21,141-142,234,334-341,
 * 496-501,802-809,955-962,1061-1068,1185-1194).  POD, passed by pointer from the host. */
typedef struct mh_config {
  int32_t family;        /* mh_family */
  int32_t easy_margin;   /* ArcFace / MagFace */
  int32_t sphere_m;      /* SphereFace integer m in 0..5 */
  int32_t plus;          /* ElasticFace plus (the re-assignment itself is done by the caller) */
  float s, m;
  float mv_weight;       /* MV_Softmax */
  float momentum;        /* CurricularFace */
  float h, t_alpha;      /* AdaFace */
  float l_margin, u_margin, l_a, u_a; /* MagFace */
  float sphere_lambda;   /* SphereFace annealing value for THIS step (host computes, criterion.py:60) */
  float reserved;
} mh_config;

/* Mutable head state living on the device: [0]=CurricularFace t, [1]=AdaFace batch_mean,
 * [2]=AdaFace batch_std, [3]=loss_g of the last forward (MagFace), [4]=the CurricularFace t used by
 * THIS forward's hard-negative modulation (criterion.py:575; equals [0] when update_state != 0). */
#define MH_STATE_FLOATS 8

const char* mh_version(void);
const char* mh_last_error(void);
/* 0 if the current device can run the kernels (compute capability 10.x), else MH_ERR_UNSUPPORTED. */
int mh_device_check(void);

/* ---- prologues (HBM-bound) ------------------------------------------------------------------ */

/* Replaces F.normalize(self.weight) / F.normalize(self.kernel, dim=0) (criterion.py:65,174,264,404,
 * 541,864,987,1096,1252).  Reads the fp32 parameter in its own layout (ld = row pitch in elements:
 * D for CD, C for DC), writes w_hat as bf16 [C_pad, 512] (rows >= C zeroed), optional fp32 copy
 * w_hat32 [C, 512] (NULL to skip) and inv_norm[C] = 1/max(|w_j|,1e-12). */
int mh_prologue_w(const float* W, int layout, int64_t C, int64_t ld, void* w_hat_bf16, int64_t C_pad,
                  float* w_hat32, float* inv_norm, void* stream);

/* Optimizer step of the class-centre parameter fused with the NEXT step's mh_prologue_w.
 * Replaces optim.SGD(..., momentum=0.9, weight_decay=5e-4).step() on the head parameter (model_utils.py:557,
 * applied at model_utils.py:186 through GradScaler.step) followed by the F.normalize of the next forward.
 * In place, per element (torch.optim.SGD with dampening 0, no nesterov; momentum_buf starts at zero):
 *   g' = grad / *grad_scale + weight_decay * w;  momentum_buf = momentum * momentum_buf + g';  w -= lr * momentum_buf
 * and, from the updated W in the same pass, w_hat bf16 [C_pad, 512] and inv_norm[C] exactly as mh_prologue_w writes
 * them.  grad and momentum_buf have W's layout and pitch.  grad_scale / found_inf are GradScaler's device scalars
 * (NULL = no scaling / never skip); a non-zero *found_inf leaves W and momentum_buf untouched (w_hat is rebuilt from
 * the unchanged W).  Algorithmic bytes per class: 3 x 2048 read + 2 x 2048 + 1024 + 4 written. */
int mh_sgd_step_w(float* W, int layout, int64_t C, int64_t ld, const float* grad, float* momentum_buf, float lr,
                  float momentum, float weight_decay, const float* grad_scale, const float* found_inf,
                  void* w_hat_bf16, int64_t C_pad, float* inv_norm, void* stream);

/* Replaces F.normalize(x) + torch.norm(x) (criterion.py:65,95,173,192,263,298,...) and the target
 * cosine gather (criterion.py:417,552).  labels are GLOBAL class ids (int64); this shard owns
 * [c_offset, c_offset + C).  Outputs: x_hat bf16 [B_pad,512] (rows >= B zeroed), x_hat32 fp32
 * [B,512], xnorm[B], t_raw[B] = <x_hat_i, w_hat_{y_i}> in fp32 (0 when the label is not owned),
 * label_local[B_pad] (int32; -1 when not owned or row >= B).  inv_norm may be NULL for the [C, 512] layout: 1/|w_y| is
 * then taken from the gathered row itself (same arithmetic as mh_prologue_w).  c_total = class count of the WHOLE head (= C when
 * unsharded): a label outside [0, c_total) sets t_raw to NaN so that the loss is NaN on every rank (NaN survives the
 * all-reduce of t_raw; the reference's one_hot.scatter_ fails on such a label, criterion.py:290-291). */
int mh_prologue_x(const void* x, int x_dtype, int64_t B, int64_t B_pad, const int64_t* labels,
                  const float* W, int layout, int64_t C, int64_t ld, int64_t c_offset,
                  const float* inv_norm, void* x_hat_bf16, float* x_hat32, float* xnorm, float* t_raw,
                  int32_t* label_local, int64_t c_total, void* stream);

/* Per-row margin terms + batch-global state updates (CurricularFace EMA criterion.py:570-573,
 * AdaFace batch statistics criterion.py:876-885, MagFace loss_g criterion.py:1248).  margins may be
 * NULL except for ElasticFace (one sampled margin per row, criterion.py:1003-1012).  state is
 * read-modify-written when update_state != 0.  rowp is [MH_RP_PLANES, ldp]. Single small kernel. */
int mh_row_params(const mh_config* cfg_host, int64_t B, const float* xnorm, const float* t_raw,
                  const float* margins, float* state, int update_state, float* rowp, int64_t ldp,
                  void* stream);

/* ---- tensor-core path (bf16 operands, fp32 accumulate, tcgen05 + TMEM + TMA) ------------------ */

/* Number of statistics records the forward keeps per row: one per CTA pair and 128-column epilogue half
 * (2 * (#SMs / 2) = 148 on a B200; independent of C_pad, which is accepted for ABI stability). */
int64_t mh_fwd_num_tiles(int64_t C_pad);

/* Host-only test hook: the static tile schedule of the A-stationary forward / backward-G kernels for `units` CTA
 * pairs, m_tiles (<= units) 256-row tiles and n_tiles 256-class tiles.  Writes up to cap (pair, m_tile, n_tile)
 * int32 triples in per-pair execution order and returns the total number of tiles scheduled (or a negative status). */
int64_t mh_tc_schedule_tiles(int units, int m_tiles, int n_tiles, int32_t* out, int64_t cap);

/* 1 when the head can use the fixed-reference softmax: fixed logit scale s with s*log2(e)*(umax+1) <= 200 (umax =
 * largest possible z/s on a non-target column: 1, MV 2w-1, Curricular 2) and C >= 2.  Then every term
 * exp2(z*log2e - ref), ref = s*log2e*umax - 102, is a normal fp32/bf16 number, no running max is needed and the row
 * sums cannot overflow.  SphereFace (scale = |x|) and large s use the online-max forward. */
int mh_tc_fixref_ok(const mh_config* cfg_host, int64_t C);
/* 1 when the forward may stash for the backward: mh_tc_fixref_ok and cos_ij recoverable from the stashed exponential
 * (u = cos, or MV-Softmax's invertible u = w*cos + w - 1; not CurricularFace's cos*(t + cos)).  Otherwise: recompute. */
int mh_tc_stash_ok(const mh_config* cfg_host, int64_t C);
/* 1 when the head fails mh_tc_stash_ok but may use the GUARDED stash of mh_step_forward (stash == 2): CurricularFace
 * (criterion.py:491-587; s*log2e*3 = 277 binades at s = 64), SphereFace (criterion.py:12-107; the scale is |x_i|) and
 * any family whose s is too large for the proof.  The fixed reference ref_i = scale_i*log2e*umax - 102 still rules out
 * overflow; whether terms flushed to zero could matter is checked per row on the device after the forward. */
int mh_tc_stash_guarded_ok(const mh_config* cfg_host, int64_t C);

/* Fused cos-GEMM + margin + softmax statistics (replaces F.linear/torch.mm at criterion.py:65,176,267,
 * 408,545,868,990,1100,1256, the elementwise margin passes and nn.CrossEntropyLoss's log-softmax,
 * model_utils.py:179).  stats_tiles is [num_tiles, MH_ST_PLANES, B_pad].  B_pad must be a multiple of 256.
 * stash_bf16 == NULL: nothing of size B x C is written (inference / recompute-backward mode).
 * stash_bf16 != NULL (training, needs mh_tc_stash_ok): additionally writes E'_ij = exp2(z_ij*log2e - ref_i) * du_ij/dcos
 * as bf16, class-tiled [C_pad/128][B_pad][128] (element (row i, class j) at ((j/128)*B_pad + i)*128 + j%128), with the
 * target column, padded classes and rows >= B set to 0.  The backward then needs no logit recompute:
 * G_ij = rho_i * E'_ij for j != y_i (see mh_stash_prep). */
int mh_tc_forward(const mh_config* cfg_host, const void* x_hat_bf16, int64_t B, int64_t B_pad,
                  const void* w_hat_bf16, int64_t C, int64_t C_pad, const float* rowp, int64_t ldp,
                  const int32_t* label_local, const float* state, float* stats_tiles, void* stash_bf16, void* stream);

/* Merged W prologue + forward (optional, [C, 512] parameters: ArcFace, SphereFace, MV-Softmax): mh_prologue_w as a ROLE of
 * the forward launch.  A few CTA pairs normalise the class centres tile by tile (fp32 W -> bf16 w_hat + inv_norm, bits
 * identical to mh_prologue_w), the other pairs run mh_tc_forward's kernel and pick every w_hat tile up from L2 right after
 * it was written: the HBM-bound prologue disappears under the tensor-bound forward and the forward's read of w_hat never
 * reaches DRAM.  Outputs as mh_prologue_w + mh_tc_forward together.  mh_prologue_x / mh_row_params must have run before
 * (mh_prologue_x with inv_norm == NULL, since inv_norm is an OUTPUT here).
 * Query: x_hat_bf16 == NULL writes 1 / 0 to *eligible_host (needs >= 8 class tiles per CTA pair, one launch of row tiles);
 * when 0, call mh_prologue_w and mh_tc_forward instead.  ready_ws: [C_pad / 256 + 1] ints of scratch (zeroed here). */
int mh_tc_forward_pw(const mh_config* cfg_host, const void* x_hat_bf16, int64_t B, int64_t B_pad, const float* W,
                     int layout, int64_t ld, void* w_hat_bf16, float* inv_norm, int64_t C, int64_t C_pad,
                     const float* rowp, int64_t ldp, const int32_t* label_local, const float* state,
                     float* stats_tiles, void* stash_bf16, int* ready_ws, int* eligible_host, void* stream);

/* Recompute backward, step 1: recompute the logit tiles and write G = (P - Y) * dz/dcos as bf16 (never the logits), in the
 * same class-tiled layout as the stash, so each 128-class slab of all rows is one contiguous block (DRAM-page friendly
 * for both consumers); mh_tc_backward_dx / _dw / _dw_fused expect this layout.  lse2 = rowout plane MH_RO_LSE2.
 * r_colsum [C_pad] (may be NULL) receives r_j = sum_i G_ij * cos_ij = w^_j . dw^_j, the projection term of
 * the normalise-backward of W; it is zeroed (stream-ordered) before the kernel accumulates into it. */
int mh_tc_backward_g(const mh_config* cfg_host, const void* x_hat_bf16, int64_t B, int64_t B_pad,
                     const void* w_hat_bf16, int64_t C, int64_t C_pad, const float* rowp, int64_t ldp,
                     const int32_t* label_local, const float* state, const float* lse2, void* G_bf16,
                     float* r_colsum, void* stream);

/* dx_hat partials = G . w_hat, split over the class dimension.
 * Returns the number of splits through *n_split_host (call with out == NULL to query).
 * out is [n_split, B_pad, 512] fp32.  sync_ws (may be NULL): MH_DX_SYNC_INTS ints of caller-owned scratch; when given
 * and every tile of the launch is resident at once, the CTAs that stream the same w^ k-blocks (one split, all row
 * tiles) rendezvous every 16 k-blocks so that w^ is fetched from HBM once per split, not once per row tile. */
#define MH_DX_SYNC_INTS 64
int mh_tc_backward_dx(const void* G_bf16, int64_t B_pad, int64_t C_pad, const void* w_hat_bf16,
                      float* out, int* n_split_host, int* sync_ws, void* stream);

/* Stash mode: the same GEMM on the stash E' (out = E' . w_hat; scale the rows by rho afterwards, mh_stash_dx_combine).
 * While the tensor cores run, the kernel's idle epilogue warps re-read every E' tile from shared memory and store
 * r_colsum[b][j] = sum_{i in 128-row block b} rho_i * E'_ij * cos_ij for b < B_pad/128 (r_colsum is [B_pad/128, C_pad];
 * plain stores, every element written once: no atomics, bit-reproducible), with
 * cos_ij = log2(E'_ij)/(s log2e) + ref/(s log2e) recovered from the stash itself: the non-target part of the projection
 * term w^_j . dw^_j that mh_tc_backward_dw_fused needs, passed there with r_parts = B_pad/128 (the target column's part
 * is added by mh_stash_dw_target).  MV-Softmax: hard negatives were stashed as w * exp2(.) with u = w*cos + w - 1; the two
 * cases occupy disjoint ranges of log2(E') for a given row threshold (rowp plane MH_RP_THR), so cos is recovered exactly. */
int mh_tc_backward_dx_stash(const mh_config* cfg_host, const void* stash_bf16, int64_t B_pad, int64_t C, int64_t C_pad,
                            const void* w_hat_bf16, const float* rho, const float* rowp, int64_t ldp, float* out,
                            float* r_colsum, int* n_split_host, int* sync_ws, void* stream);

/* dw_hat = G^T . x_hat, [C_pad, 512] fp32 (unscaled, unprojected). */
int mh_tc_backward_dw(const void* G_bf16, int64_t B_pad, int64_t C_pad, const void* x_hat_bf16,
                      float* dw_hat, void* stream);

/* dw_hat = G^T . x_hat fused with the normalise-backward of W (autograd of F.normalize(self.weight), criterion.py:264):
 * dW_j = gscal[0] * (dw^_j - w^_j * r_j) / |w_j| written straight into the parameter layout (ld = row pitch), with
 * r_j = sum_{p < r_parts} r_colsum[p * C_pad + j] (1 plane from mh_tc_backward_g, B_pad/128 from mh_tc_backward_dx_stash);
 * replaces mh_tc_backward_dw + mh_norm_backward_w and their 4*C*d-byte dw_hat round trip.  x_hat_bf16 is x^ (recompute
 * mode, G from mh_tc_backward_g) or the rho-scaled rows of mh_stash_prep (stash mode, G = the stash). */
int mh_tc_backward_dw_fused(const void* G_bf16, int64_t B_pad, int64_t C, int64_t C_pad, const void* x_hat_bf16,
                            const void* w_hat_bf16, const float* inv_norm, const float* r_colsum, int r_parts,
                            const float* gscal, int layout, float* dW, int64_t ld, void* stream);

/* dW with the projection term taken from the accumulators: r_j = w^_j . dw^_j is formed inside the kernel (the two
 * CTA pairs holding the two d halves of a class tile exchange four partial dots per class through rpart_ws, summed in a
 * fixed order: bit-reproducible, no atomics on data), so no r_colsum input and no side pass in the dx kernel.
 * rpart_ws: [4 * C_pad] floats, flag_ws: [C_pad / 128] ints - caller-owned scratch (flag_ws is zeroed here). */
int mh_tc_backward_dw_proj(const void* G_bf16, int64_t B_pad, int64_t C, int64_t C_pad, const void* x_hat_bf16,
                           const void* w_hat_bf16, const float* inv_norm, const float* gscal, int layout,
                           float* dW, int64_t ld, float* rpart_ws, int* flag_ws, void* stream);

/* Merged backward (optional): the dx^ GEMM and the self-projecting dW GEMM in ONE persistent kernel.  The first
 * m_tiles * n_split CTA pairs compute the dx^ partials over interleaved 256-class chunks, the other pairs compute dW; both
 * roles walk the classes in the same direction and throttle each other to stay within 48 class tiles, so the B x C buffer
 * (G or the stash) and w^ are fetched from HBM once and hit in L2 for the other role (6.15 GB less DRAM traffic per step
 * at cfg4 than mh_tc_backward_dx + mh_tc_backward_dw_*).  No r_colsum, no dx side pass (dW projects itself).
 * Query: dxhat_part == NULL writes the split count to *n_split_host - 0 when the shape is not eligible (fewer than 8
 * class tiles per pair, or more row tiles than half the pairs): run the two kernels back to back then.
 * dxhat_part [n_split, B_pad, 512]; rpart_ws [4 * C_pad] floats, flag_ws [C_pad / 128] ints, prog_ws [2] ints (scratch,
 * zeroed here).  xs_bf16: x^ (recompute mode) or rho * x^ (stash mode), as for mh_tc_backward_dw_fused. */
int mh_tc_backward_dxdw(const void* G_bf16, int64_t B_pad, int64_t C, int64_t C_pad, const void* w_hat_bf16,
                        const void* xs_bf16, const float* inv_norm, const float* gscal, int layout, float* dW,
                        int64_t ld, float* dxhat_part, int* n_split_host, float* rpart_ws, int* flag_ws,
                        int* prog_ws, void* stream);

/* ---- stash backward: O(B*d) helpers (see mh_tc_forward's stash) -------------------------------------- */

/* rho_i = scale_i * 2^(ref_i - lse2_i), gty_i = G_{i,y_i} = (P_iy - 1) * dz_iy/dcos (rowout AUX0 * rowp DZT) and
 * xs = bf16(rho_i * x^_i) [B_pad, 512] (rows >= B zero): the B operand of mh_tc_backward_dw_fused in stash mode. */
int mh_stash_prep(const mh_config* cfg_host, const float* rowp, int64_t ldp, const float* rowout, int64_t ldo,
                  const float* x_hat32, int64_t B, int64_t B_pad, void* xs_bf16, float* rho, float* gty, void* stream);

/* dxhat [B, 512] = rho_i * sum_splits dxhat_part + gty_i * w^_{y_i} (target term only where label_local >= 0).
 * rho == NULL: the plain sum of the split-K partials (gty / label_local / w_hat ignored; recompute mode, sharded). */
int mh_stash_dx_combine(const float* dxhat_part, int n_split, int64_t split_stride, const float* rho, const float* gty,
                        const int32_t* label_local, const void* w_hat_bf16, int64_t B, float* dxhat, void* stream);

/* Adds the target-column term to dW after mh_tc_backward_dw_fused ran on the stash:
 * dW_j += gscal[0]/|w_j| * (delta_j - w^_j (w^_j . delta_j)), delta_j = sum_{i: y_i = j} gty_i x^_i.  Deterministic. */
int mh_stash_dw_target(const float* gty, const int32_t* label_local, const float* x_hat32, const void* w_hat_bf16,
                       const float* inv_norm, const float* gscal, int64_t B, int layout, float* dW, int64_t ld,
                       void* stream);

/* ---- exact fp32 path (SIMT; materialises S = x^ w^T [B, C]; small C, tests, compat mode) ------- */

/* C[M,N] (ldc) = A[M,K] . B[K,N] with arbitrary element strides (row, col) for A and B. fp32 FMA. */
int mh_sgemm_strided(int64_t M, int64_t N, int64_t K, const float* A, int64_t a_rs, int64_t a_cs,
                     const float* B, int64_t b_rs, int64_t b_cs, float* C, int64_t ldc, void* stream);

/* From dense cosines S [B, C] (ld = lds_): per-row statistics as ONE tile (stats [MH_ST_PLANES,
 * B_pad]) and, when non-NULL, the materialised reference outputs pre [B,C] and logits [B,C]
 * (criterion.py:197,301,... the 4-tuple's first element). */
int mh_dense_forward(const mh_config* cfg_host, const float* S, int64_t lds_, int64_t B, int64_t B_pad,
                     int64_t C, const float* rowp, int64_t ldp, const int32_t* label_local,
                     const float* state, float* stats, float* pre, float* logits, void* stream);

/* In-place S -> dcos.  If dlogits == NULL: dcos = (P - Y) * dz/dcos using lse2 (fused CE).
 * Otherwise dcos = dlogits * dz/dcos + dpre * dpre/dcos (compat mode: caller-side loss on the
 * materialised logits) and rowaux [2, B] receives the AUX0/AUX1 row terms of that upstream grad. */
int mh_dense_backward_dc(const mh_config* cfg_host, float* S, int64_t lds_, int64_t B, int64_t C,
                         const float* rowp, int64_t ldp, const int32_t* label_local, const float* state,
                         const float* lse2, const float* dlogits, const float* dpre, float* rowaux,
                         void* stream);

/* ---- reductions / finalisers -------------------------------------------------------------------- */

/* Merge per-tile (or per-shard) statistics: stats_in [n_parts, MH_ST_PLANES, lds_] ->
 * stats_out [MH_ST_PLANES, lds_] (online-softmax merge: max, rescaled sums, counts add). */
int mh_merge_stats(const float* stats_in, int64_t n_parts, int64_t B, int64_t lds_, float* scratch,
                   float* stats_out, void* stream);
/* scratch: [MH_MERGE_BLOCKS, MH_ST_PLANES, lds_] floats */
#define MH_MERGE_BLOCKS 64

/* Final per-row results from merged statistics + scalars:
 * rowout [MH_RO_PLANES, ldo]; scalars[0]=mean CE loss, [1]=acc@1 %, [2]=acc@5 % (metrics.py:3-16),
 * computed over rows [0,B) with divisor B_total (global batch, >= B); scalars[3] = state[3] (loss_g of this forward;
 * 0 when state is NULL). */
int mh_finalize_rows(const float* stats, int64_t lds_, const float* rowp, int64_t ldp, int64_t B,
                     int64_t B_total, int sphere, float* rowout, int64_t ldo, float* scalars, const float* state,
                     void* stream);

/* gscal[0] = g_loss[0] / B_total, gscal[1] = g_lossg[0] (NULL pointers read as 0): the device-side upstream-gradient
 * scalars every backward kernel takes, so that a GradScaler-scaled backward (model_utils.py:185) needs no host sync. */
int mh_make_gscal(const float* g_loss, const float* g_lossg, int64_t B_total, float* gscal, void* stream);

/* dx = (dxh - x^ (x^.dxh)) / |x| + dn * x^   with dxh = gz * sum_splits dxhat and
 * dn = gz * (AUX0 * DZT_DN + AUX1) + g_lossg * DLG_DN  (autograd of F.normalize + the |x| paths of
 * SphereFace criterion.py:95-105 and MagFace criterion.py:1244-1266).  dxhat is [n_split] partials
 * of [*, 512] fp32 at stride split_stride floats; aux0/aux1 are the AUX planes (rowout or rowaux).
 * gscal = {gz, g_lossg} on the device.  dx is written in x_dtype. */
int mh_norm_backward_x(const float* dxhat, int n_split, int64_t split_stride, const float* x_hat32,
                       const float* xnorm, const float* rowp, int64_t ldp, const float* aux0,
                       const float* aux1, const float* gscal, int64_t B, void* dx, int x_dtype,
                       void* stream);

/* dW_j = g * k_j * (dw^_j - w^_j (w^_j . dw^_j)) / |w_j|, written in the parameter's own layout
 * (ld = row pitch of dW).  w_hat given as bf16 [C_pad,512] or fp32 [C,512] (exactly one non-NULL).
 * class_scale k [C] may be NULL (= 1): VPL-ArcFace passes 1 - a_j (criterion.py:724). */
int mh_norm_backward_w(const float* dw_hat, const void* w_hat_bf16, const float* w_hat32,
                       const float* inv_norm, const float* gscal, const float* class_scale, int64_t C, int layout,
                       float* dW, int64_t ld, void* stream);

/* gscal[0] = upstream_grad(loss_id) / B_total, gscal[1] = upstream_grad(loss_g): device floats so
 * that a GradScaler-scaled backward (model_utils.py:185) needs no host sync. */

/* ---- whole-phase entry points (single GPU, tensor-core path) -------------------------------------------------- */

/* Workspace descriptor shared by mh_step_forward / mh_step_backward: every pointer is a caller-owned DEVICE buffer with
 * the shape given in DESIGN.md section 3 (the same buffers the per-kernel entry points take). */
typedef struct mh_step_ws {
  int64_t B, B_pad, C, C_pad;  /* B_pad, C_pad: multiples of 256 */
  int64_t ld;                  /* row pitch of W and dW in elements (512 for [C,D], C for [D,C]) */
  int32_t layout;              /* mh_layout of W / dW */
  int32_t x_dtype;             /* mh_dtype of x and dx */
  void* w_hat;                 /* bf16 [C_pad, 512] */
  float* inv_norm;             /* [C] */
  void* x_hat;                 /* bf16 [B_pad, 512] */
  float* x_hat32;              /* [B, 512] */
  float* xnorm;                /* [B] */
  float* t_raw;                /* [B] */
  int32_t* label_local;        /* [B_pad] */
  float* rowp;                 /* [MH_RP_PLANES, B_pad] */
  float* stats_tiles;          /* [n_tiles, MH_ST_PLANES, B_pad] */
  int64_t n_tiles;             /* = mh_fwd_num_tiles(C_pad) */
  float* merge_scratch;        /* [MH_MERGE_BLOCKS, MH_ST_PLANES, B_pad] */
  float* stats;                /* [MH_ST_PLANES, B_pad] */
  float* rowout;               /* [MH_RO_PLANES, B_pad] */
  void* bc;                    /* bf16 [B_pad, C_pad] class-tiled: the stash or G (NULL: forward-only use) */
  void* xs;                    /* bf16 [B_pad, 512] (stash mode) */
  float* rho;                  /* [B_pad] (stash mode) */
  float* gty;                  /* [B_pad] (stash mode) */
  float* dxhat_part;           /* [part_splits, B_pad, 512] */
  int64_t part_splits;         /* >= the split count mh_tc_backward_dx reports for (B_pad, C_pad) */
  float* dxhat_full;           /* [B_pad, 512] (stash mode) */
  float* gscal;                /* [2] */
  float* r_colsum;             /* [B_pad / 128, C_pad]: projection partials from the dx side pass / backward-G sums; NULL selects
                                  the self-projecting dW kernel (mh_tc_backward_dw_proj), which needs the next two instead */
  float* rpart;                /* [4, C_pad] or NULL */
  int32_t* rflag;              /* [C_pad / 128] or NULL */
  int32_t* dx_sync;            /* [MH_DX_SYNC_INTS] or NULL */
  int32_t* pw_ready;           /* [C_pad / 256 + 1] or NULL; non-NULL selects the merged prologue + forward kernel
                                  (mh_tc_forward_pw) whenever the W prologue has to run and the head / shape is eligible */
  int32_t* prog;               /* [2] or NULL; non-NULL (with rpart / rflag) selects the merged dx + dW kernel
                                  (mh_tc_backward_dxdw) whenever both gradients are wanted and the shape is eligible */
  int32_t* guard;              /* [1] or NULL; needed by the guarded stash (stash == 2): 1 after a forward whose fixed-reference
                                  sums could not be trusted and that therefore re-ran the general path on the device */
  void* graph_cache;           /* handle from mh_step_cache_create or NULL: a phase is captured into a CUDA graph per distinct
                                  argument set and replayed whenever that set comes back (see below) */
} mh_step_ws;

/* CUDA-graph cache of the two phases (HOST object, tied to the device that is current at creation).  mh_step_forward /
 * mh_step_backward called with ws->graph_cache set capture a phase the first time they see a set of arguments
 * (hyper-parameters, workspace descriptor, every pointer and flag, byte by byte) and replay it with one cudaGraphLaunch on
 * the caller's stream whenever the same set comes back; the capture happens on a private stream of the cache, so the
 * caller's stream may be the legacy default stream.  A few keys are kept (LRU); a phase whose key changes on every call
 * (fresh addresses, SphereFace's annealed lambda) backs off to plain launches.  A caller
 * whose stream is itself being captured gets plain launches.  Same kernels, arguments and order as without the cache:
 * bit-identical results.  Thread-safe; destroy only when no call is in flight.  MH_STEP_GRAPH=0 in the environment
 * disables replay. */
int mh_step_cache_create(void** cache_out);
int mh_step_cache_destroy(void* cache);
/* counts3 = {phases replayed from a graph, phases captured (and then launched as a graph), phases issued as plain
 * launches while backing off}; phases of a caller that is itself capturing are not counted. */
int mh_step_cache_stats(void* cache, int64_t* counts3);

/* Forward of one step: mh_prologue_w (skipped when run_prologue_w == 0: w_hat / inv_norm already hold this W, e.g.
 * after mh_sgd_step_w), mh_prologue_x, mh_row_params, mh_tc_forward (stashing into ws->bc when stash != 0; stash == 1
 * needs mh_tc_stash_ok), mh_merge_stats, mh_finalize_rows.  scalars[4] = {mean CE loss, acc@1 %, acc@5 %, loss_g}.
 * stash == 2, the GUARDED stash (needs mh_tc_stash_guarded_ok and ws->guard; CurricularFace criterion.py:491-587,
 * SphereFace criterion.py:12-107, any family at s > 69): the fixed-reference forward + stash runs speculatively, the
 * finaliser checks on the device that no row's sum can have lost anything to underflow (row sum >= C 2^-102) and
 * writes ws->guard; the general online-max forward, its merge and its finaliser are enqueued behind it as launches
 * that return at once unless the flag is set.  No host synchronisation, same results as the general path to 2^-24.
 * Replaces the head forward + nn.CrossEntropyLoss + accuracy of model_utils.py:177-182 in ONE host call. */
int mh_step_forward(const mh_config* cfg_host, const mh_step_ws* ws, const void* x, const int64_t* labels,
                    const float* W, const float* margins, float* state, int update_state, int run_prologue_w,
                    int stash, float* scalars, void* stream);

/* Backward of the step whose forward just ran on the same workspace (stash must match): mh_make_gscal, then either the
 * stash backward (mh_stash_prep, mh_tc_backward_dx_stash, mh_stash_dx_combine, mh_norm_backward_x,
 * mh_tc_backward_dw_fused, mh_stash_dw_target) or the recompute backward (mh_tc_backward_g, mh_tc_backward_dx,
 * mh_norm_backward_x, mh_tc_backward_dw_fused); with ws->r_colsum == NULL the dW kernel is mh_tc_backward_dw_proj
 * and the dx GEMM runs without its side pass.  g_loss / g_lossg: device scalars (upstream gradients of loss / loss_g; NULL = 0).
 * dx ([B, 512] in x_dtype) and dW (parameter layout, fp32) may each be NULL to skip that gradient.
 * stash == 2 (guarded stash): the stash backward with the self-projecting dW kernel (ws->r_colsum must be NULL); a
 * gated mh_tc_backward_g first rewrites ws->bc with the recomputed G when the forward raised ws->guard, and
 * mh_stash_prep then uses rho = 1 and no separate target-column term.
 * Replaces autograd's backward of model_utils.py:185 through the head in ONE host call. */
int mh_step_backward(const mh_config* cfg_host, const mh_step_ws* ws, int stash, const float* state,
                     const float* g_loss, const float* g_lossg, void* dx, float* dW, void* stream);

/* ---- gated / guarded forms (the building blocks of the guarded stash, for drivers that sequence the kernels themselves:
 * the class-sharded head puts collectives between them) ---------------------------------------------------------------
 * gate: device flag read by every thread of the launch before anything else; the launch does nothing unless
 * (*gate != 0) == (gate_on != 0).  gate == NULL: always runs (then each function equals its plain namesake). */

/* mh_tc_forward with stash_kind: 0 no stash (stash_bf16 NULL), 1 the proven stash (mh_tc_stash_ok), 2 the guarded stash
 * (mh_tc_stash_guarded_ok: fixed-reference sums + stash whose validity the caller checks with mh_finalize_rows_ex). */
int mh_tc_forward_ex(const mh_config* cfg_host, const void* x_hat_bf16, int64_t B, int64_t B_pad, const void* w_hat_bf16,
                     int64_t C, int64_t C_pad, const float* rowp, int64_t ldp, const int32_t* label_local,
                     const float* state, float* stats_tiles, void* stash_bf16, int stash_kind, const int32_t* gate,
                     int gate_on, void* stream);
/* mh_tc_backward_g behind a gate (r_colsum must be NULL when gated). */
int mh_tc_backward_g_ex(const mh_config* cfg_host, const void* x_hat_bf16, int64_t B, int64_t B_pad,
                        const void* w_hat_bf16, int64_t C, int64_t C_pad, const float* rowp, int64_t ldp,
                        const int32_t* label_local, const float* state, const float* lse2, void* G_bf16,
                        float* r_colsum, const int32_t* gate, int gate_on, void* stream);
/* mh_merge_stats behind a gate. */
int mh_merge_stats_ex(const float* stats_in, int64_t n_parts, int64_t B, int64_t lds, float* scratch, float* stats_out,
                      const int32_t* gate, int gate_on, void* stream);
/* mh_finalize_rows behind a gate, and / or with the guard of the guarded stash: guard_flag != NULL also writes
 * *guard_flag = 1 if any row's fixed-reference sum is below guard_min_l (= C_total * 2^-102: everything that can have been
 * flushed to zero is then < 2^-24 of the sum), else 0. */
int mh_finalize_rows_ex(const float* stats, int64_t lds, const float* rowp, int64_t ldp, int64_t B, int64_t B_total,
                        int sphere, float* rowout, int64_t ldo, float* scalars, const float* state, float guard_min_l,
                        int32_t* guard_flag, const int32_t* gate, int gate_on, void* stream);
/* mh_stash_prep for the guarded stash: with *fallback != 0 (the forward fell back and mh_tc_backward_g_ex rewrote the
 * B x C buffer with G itself) rho = 1, gty = 0 and xs = bf16(x^). */
int mh_stash_prep_ex(const mh_config* cfg_host, const float* rowp, int64_t ldp, const float* rowout, int64_t ldo,
                     const float* x_hat32, int64_t B, int64_t B_pad, void* xs_bf16, float* rho, float* gty,
                     const int32_t* fallback, void* stream);

/* ---- VPL-ArcFace (criterion.py:619-762) ------------------------------------------------------------------ */

/* Mixed class vectors of the virtual-proxy head: with a_j = lamda * 1[life_j > 0] (life AFTER this step's decay,
 * criterion.py:716-717), every non-target cosine of the reference is x^_i . v_j with
 *     v_j = (1 - a_j) * w^_j + a_j * mem_j / max(|mem_j|, 1e-12)                       (criterion.py:720-724)
 * so the B x C GEMM runs on v (bf16 [C_pad, 512], rows >= C zero) instead of w^.  alpha_out[j] = a_j (fp32 [C]); the
 * interpolation weights are formed in fp32 exactly as the reference's float32 active_mask does.  One warp per class;
 * HBM-bound: 2*d (w^ bf16) + 4*d (mem fp32) bytes read, 2*d written per class. */
int mh_vpl_mix(const void* w_hat_bf16, const float* mem, const float* life, float lamda, int64_t C, int64_t C_pad,
               void* v_bf16, float* alpha_out, void* stream);

/* ---- pair verification (the device-side piece of the LFW evaluator) ------------------- */

/* cos_out[i] = <e1_i, e2_i> / (max(|e1_i|,1e-12) * max(|e2_i|,1e-12)) for N embedding pairs of dimension d
 * (row pitches ld1, ld2 in elements; dtype MH_F32 / MH_BF16 / MH_F16): F.normalize on both backbone outputs and the
 * row-wise dot of model_utils.py:333-335, 367-369, 392-394.  HBM-bound, one warp per pair. */
int mh_pair_cosine(const void* e1, const void* e2, int dtype, int64_t N, int64_t d, int64_t ld1, int64_t ld2,
                   float* cos_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MARGIN_HEAD_H_ */
