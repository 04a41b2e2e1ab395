/* scv.h — C ABI of libscv.so, the sm_100a kernel library behind scrubvae_b200.
 *
 * The reference (tdunnlab/scrubvae) has no FFI: its hot path is PyTorch library calls.  Each entry
 * point below replaces the library call sites of one part of the SC-VAE training step; the
 * reference file:line it replaces is cited per function (paths relative to
 * /root/reference/src/scrubvae).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - plain pointers and sizes only; every struct field is 8 bytes wide (pointer, int64_t, double)
 *    so that the ctypes mirror cannot mis-pad;
 *  - all device buffers are owned by the caller (PyTorch caching allocator); the library never
 *    allocates, frees or synchronises; every call only enqueues work on `stream` (a cudaStream_t);
 *  - return value 0 = ok, >0 = cudaError_t, <0 = argument error; scv_last_error() gives the text;
 *  - activations are fp32, channels-last, halo-padded:  X[b][halo + l][c], halo rows are zero.
 *    A convolution row (b,l) is then a CONTIGUOUS run of k*C floats, so every conv / transposed
 *    conv / linear, forward, dgrad and wgrad, is one "overlapping-row GEMM" (scv_gemm / scv_wgrad).
 */
#ifndef SCV_H_
#define SCV_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCV_ACT_NONE 0
#define SCV_ACT_RELU 1
#define SCV_ACT_TANH 2
#define SCV_ACT_RELUMASK 3 /* y = (R > 0) ? y : 0 ; R is a mask, not added */
#define SCV_ACT_ROUND_TF32 16 /* OR-ed into act: round the stored Y to TF32 (it feeds a tensor-core GEMM) */
#define SCV_ACT_ACCUM 32 /* OR-ed into act: Y += result instead of Y = result (activation NONE, no R, no stats).  The
                          * tensor-core path then cuts a long reduction over several CTAs (split-K) and adds the partial
                          * tiles with red.global.add: the caller zeroes Y first. */

/* `flags` arguments of the kernels that PRODUCE tensor-core GEMM operands: round what is stored to TF32
 * (round-to-nearest, ties away: cvt.rna.tf32.f32).  tcgen05 kind::tf32 truncates fp32 operands; rounding at
 * the producer makes the operand error unbiased.  0 keeps plain fp32 (the SCV_PREC_FP32 path). */
#define SCV_F_ROUND_TF32 1
#define SCV_F_OUT_BF16 2 /* the produced operand buffer holds bf16 elements (pointer typed float* in the prototype; strides
                          * and offsets count ELEMENTS): what SCV_PREC_BF16 GEMMs read as A / dY */
#define SCV_GATHER_SKIP_NEG 1
#define SCV_GATHER_ROUND_TF32 2
#define SCV_GATHER_OUT_BF16 4 /* dst holds bf16 elements (packed bf16 weight matrices) */
#define SCV_MODE_ROUND_TF32 8 /* scv_bnact_*: mode bit3 */
#define SCV_MODE_OUT_BF16 16  /* scv_bnact_*: mode bit4: H / U (forward) or dX (backward) hold bf16 elements */

#define SCV_PREC_FP32 0 /* FFMA, fp32 exact */
#define SCV_PREC_TF32 1 /* tcgen05 kind::tf32, fp32 accumulate in TMEM */
#define SCV_PREC_BF16 2 /* tcgen05 kind::f16: A and W (scv_gemm) / A and dY (scv_wgrad) point to bf16 elements, their strides
                         * and K count elements; fp32 accumulate in TMEM; Y, R, bias, dW, dbias stay fp32.  No FFMA
                         * fallback: a shape the tensor-core path declines is an error. */

int scv_version(void);
const char* scv_last_error(void);
/* number of kernels this library has launched in this process (bench.py: gpu_launches) */
int64_t scv_launch_count(void);

/* ---- overlapping-row GEMM ---------------------------------------------------------------
 * Y[b][l][n] = act( out_scale * sum_k A[b*a_bs + l*a_ls + k] * W[n*K + k] + bias[n % bias_mod] + R[b][l][n] )
 * for b < B, l < Lo, n < N (row l == Lo-1 only for n < n_last).  Optional per-column sums of the
 * pre-activation value and its square are accumulated into stats[0..N) and stats[N..2N) (double).
 * Replaces cuDNN conv fprop/dgrad and cuBLAS linear: model/residual.py:78-119 (ResidualBlock),
 * :136-180 (ResidualBlockTranspose), :198,219-222,264,286 (conv_in, fc_mu/fc_sigma, fc_in,
 * conv_out), model/disentangle.py:583-632 (MLPEnsemble).
 */
typedef struct {
  const float* A; int64_t a_bs, a_ls;
  int64_t B, Lo, K, N;
  const float* W;                       /* packed [N][K], K contiguous */
  const float* bias; int64_t bias_mod, bias_n; /* bias[n % bias_mod] for n < bias_n; NULL = none */
  float* Y; int64_t y_bs, y_ls, n_last;
  const float* R; int64_t r_bs, r_ls;   /* residual (or mask), same (b,l,n) addressing as Y; NULL = none */
  int64_t act; double out_scale;
  double* stats;                        /* [2][N] accumulators or NULL */
  int64_t precision;
  /* Optional fused BatchNorm(+PReLU) BACKWARD REDUCTION (bnr_sums != NULL; activation NONE, no stats / ROUND / ACCUM):
   * the stored Y(b,l,n) is the gradient g w.r.t. the output of a BN(+PReLU) layer (model/residual.py:88-89,112-119,
   * 146-147,172-180) whose pre-normalisation input is bnr_x, addressed like Y with (bnr_bs, bnr_ls), channel n % bnr_c.
   * With y = scale x + shift (bnr_chan = [scale | shift | mean | rstd] x bnr_c, as scv_bnact_fwd writes to chan_out;
   * NULL = no BatchNorm) and g' = (y < 0 ? slope : 1) g (bnr_slope NULL = no PReLU):
   *   bnr_sums[c] += sum g',  bnr_sums[C + c] += sum g' (x - mean) rstd,  bnr_sums[2C] += sum over y < 0 of g y
   * i.e. exactly what scv_bnact_bwd_reduce accumulates, without its extra pass over X and Y. */
  const float* bnr_x; int64_t bnr_bs, bnr_ls;
  const float* bnr_chan; const float* bnr_slope; int64_t bnr_c;
  double* bnr_sums;
} scv_gemm_t;
int scv_gemm(const scv_gemm_t* p, void* stream);

/* dW[n][k] += sum_{b,l} dY[b*y_bs + l*y_ls + n] * A[b*a_bs + l*a_ls + k];
 * dbias[n % bias_mod] += sum_{b,l} dY[..n] (n < bias_n).  dY must hold zeros where invalid.
 * The tensor-core path fetches whole 32-float column slabs: A and dY must be READABLE up to the next multiple of
 * 32 floats past the end of their last row (what is read there never reaches a stored result).
 * Replaces cuDNN wgrad / cuBLAS for the same layers (autograd of the sites above). */
typedef struct {
  const float* A; int64_t a_bs, a_ls;
  int64_t B, Lo, K, N;
  const float* dY; int64_t y_bs, y_ls;
  float* dW;
  float* dbias; int64_t bias_mod, bias_n;
  int64_t precision;
} scv_wgrad_t;
int scv_wgrad(const scv_wgrad_t* p, void* stream);

/* Up to 6 small INDEPENDENT problems (no output of one is an input of another) in ONE launch on the fp32 FFMA
 * path; `precision` is ignored.  Used for the four MLPs of a scrubber-head ensemble (model/disentangle.py:583-632):
 * their same-depth layers run side by side instead of as 11 + 22 serial launches. */
int scv_gemm_group(const scv_gemm_t* p, int64_t n, void* stream);
int scv_wgrad_group(const scv_wgrad_t* p, int64_t n, void* stream);

/* ---- input pack: ResVAE.encode model/residual.py:438-451 + normalize_root :428-431 --------
 * out[b][halo+w][0..nx) = x6d[b][w][:], [nx..nx+3) = 2*(root-a0)/(a1-a0)-1, rest 0 (C floats/row) */
int scv_pack_input(const float* x6d, const float* root, const float* arena, float* out,
                   int64_t B, int64_t W, int64_t nx, int64_t C, int64_t halo, int64_t flags, void* stream);

/* ---- BatchNorm1d(train/eval) + PReLU (+ x2 linear upsample) -------------------------------
 * model/residual.py:88-89,112-113,146-147,160,172-174,199.  X rows (b,l) of C floats.
 * mode bit3: round H/U (forward) or dX (backward) to TF32.
 * mode bit0: batch-norm present, bit1: PReLU present, bit2: training (batch statistics from
 * `stats` = column sums / sums of squares over `fold` column groups, `count` elements per channel;
 * running stats updated with `momentum`), else running statistics are used.
 * H gets the activated rows; U (optional) gets the 2L rows of nn.Upsample(2,'linear'). */
typedef struct {
  const float* X; int64_t x_bs, x_ls;
  int64_t B, L, C;
  const double* stats; int64_t fold; double count, eps, momentum;
  const float* gamma; const float* beta; float* running_mean; float* running_var;
  const float* slope;
  float* H; int64_t h_bs, h_ls;
  float* U; int64_t u_bs, u_ls;
  int64_t mode;
  float* chan_out; /* optional [4][C]: per-channel scale, shift, mean, rstd of this normalisation (for scv_gemm's bnr_chan) */
} scv_bnact_t;
int scv_bnact_fwd(const scv_bnact_t* p, void* stream);

/* backward of the above.  dO rows (b,l) are the gradient w.r.t. H; dU (optional) w.r.t. U.
 * reduce: sums[0..C) += sum g, sums[C..2C) += sum g*xhat, sums[2C] += dslope   (double)
 * apply : dX rows = BN backward; dgamma/dbeta/dslope (+=) param grads from `sums`. */
typedef struct {
  const float* X; int64_t x_bs, x_ls;
  int64_t B, L, C;
  const double* stats; int64_t fold; double count, eps;
  const float* gamma; const float* beta; const float* slope;
  const float* dO; int64_t o_bs, o_ls;
  const float* dU; int64_t u_bs, u_ls;
  double* sums;
  float* dX; int64_t d_bs, d_ls;
  float* dgamma; float* dbeta; float* dslope;
  int64_t mode;
  const float* chan; /* optional [4][C] table from scv_bnact_fwd (chan_out): used instead of recomputing from stats */
} scv_bnact_bwd_t;
int scv_bnact_bwd_reduce(const scv_bnact_bwd_t* p, void* stream);
int scv_bnact_bwd_apply(const scv_bnact_bwd_t* p, void* stream);

/* ---- CholeskyL + reparameterisation: model/residual.py:39-68, :305-316 ---------------------
 * ms rows: [mu (z) | sigma raw (z(z+1)/2)] with row stride ms_ld.  Writes mu (B,z), dense L (B,z,z),
 * zc rows [z | var | 0-pad] (row stride zc_ld; var = B x nvar, may be NULL).  eps NULL => z = mu. */
int scv_reparam_fwd(const float* ms, int64_t ms_ld, const float* eps, const float* var, int64_t nvar,
                    float* mu, float* L, float* zc, int64_t zc_ld, int64_t B, int64_t z, int64_t flags,
                    void* stream); /* flags: zc is rounded */
/* dms = backward of the above given dmu (B,z), dz rows (stride dz_ld), dL dense (each may be NULL);
 * dmu2 (optional, B x z) is added to dmu scaled by dmu2_scale (gradient reversal: -alpha). */
int scv_reparam_bwd(const float* ms, int64_t ms_ld, const float* eps, const float* dmu, const float* dmu2,
                    double dmu2_scale, const float* dz, int64_t dz_ld, const float* dL, float* dms,
                    int64_t dms_ld, int64_t B, int64_t z, int64_t flags, void* stream); /* flags: dms is rounded */

/* ---- prior_loss train/losses.py:138-146: loss[0] += KL/B (double, if loss != NULL);
 * if dmu/dL != NULL they get gscale[0] * dKL/dmu, dKL/dL (gscale: device scalar, NULL = 1) ------*/
int scv_kl(const float* mu, const float* L, double* loss, const float* gscale, float* dmu, float* dL,
           int64_t B, int64_t z, void* stream);

/* ---- reconstruction losses on the decoder output xh (B*W rows of ld floats: nx 6-D channels
 * then 3 normalised root channels, all after tanh):
 *   jpe  : mpjpe_loss train/losses.py:148-171 -> fwd_kin_cont6d_torch data/dataset.py:83-116 ->
 *          cont6d_to_matrix data/quaternion.py:337-353;  loss[0] += sum (target-pose)^2/(B*3*J)
 *   root : train/losses.py:216-219 with inv_normalize_root model/residual.py:433-436;
 *          loss[1] += sum (root_hat-root)^2 / B;  root_hat (F,3) is written out.
 * dxh rows get the UNIT gradients d jpe/d xh[..nx) and d root/d xh[nx..nx+3) (pad columns 0).
 * tree = [n_chains, len_0, joints_0..., len_1, ...] (int32, device).  tree_kind: 0 = both kernels are launched and decide on the
 * device which one serves this skeleton (the other exits at once); 1 = the caller knows the lane-per-chain kernel serves it
 * (<= 8 chains of <= 5 joints, every chain starting at joint 0 or at a joint an earlier chain placed, J <= 32); 2 = generic. */
int scv_recon_loss(const float* xh, int64_t ld, const float* offsets, const float* target,
                   const float* root, const float* arena, const int32_t* tree, int64_t n_tree,
                   double* loss, float* root_hat, float* dxh, int64_t F, int64_t B, int64_t J,
                   int64_t tree_kind, void* stream);

/* draw[b][halo+w][c] = (g_jpe*dxh[c<nx] | g_root*dxh[nx<=c<nx+3]) * (1 - xh^2); g_* are device
 * scalars (NULL = 0).  Backward of tanh at model/residual.py:291 fused with the loss scales. */
int scv_out_bwd(const float* xh, const float* dxh, int64_t ld, const float* g_jpe, const float* g_root,
                int64_t nx, float* draw, int64_t d_bs, int64_t d_ls, int64_t B, int64_t W, int64_t flags,
                void* stream);

/* ---- gradient-reversal head loss train/losses.py:267-284 (nested normalisation, quirk i) ------
 * pred[e] (B x d, row stride ld) for e < n_ens; target (B x d float) or labels (B int64, CE).
 * loss[0] += sum_e w_e * L_e with w_e = c^-(n_ens-e), c = n_ens*num_keys*B (if loss != NULL);
 * dpred[e] (if dpred != NULL) = gscale[0] * w_e * dL_e/dpred (gscale: device scalar, NULL = 1). */
int scv_gr_loss(const float* const* pred, float* const* dpred, int64_t ld, int64_t n_ens,
                const float* target, const int64_t* labels, int64_t B, int64_t d, int64_t num_keys,
                double* loss, const float* gscale, void* stream);

/* ---- gather  dst[i] = idx[i] >= 0 ? src[idx[i]] : 0  (weight repack / gradient unpack;
 * flags & SCV_GATHER_SKIP_NEG leaves dst[i] untouched where idx[i] < 0, & SCV_GATHER_ROUND_TF32 rounds) */
int scv_gather(const float* src, const int32_t* idx, float* dst, int64_t n, int64_t flags, void* stream);

/* ---- step tail train/trainer.py:160-165: clip_grad_norm_(max_norm) + Adam/AdamW/SGD ----------
 * sumsq[0] += sum g^2 (double).  The update kernel derives the clip coefficient
 * min(1, max_norm/(sqrt(sumsq)*gscale + 1e-6)) itself; gscale scales grads first (1/world). */
int scv_sumsq(const float* g, int64_t n, double* sumsq, void* stream);
typedef struct {
  float* p; const float* g; float* m; float* v; int64_t n;
  const double* sumsq; double max_norm, gscale;
  double lr, beta1, beta2, eps, weight_decay; int64_t step;
  const double* hyper; /* optional device [lr, step]: overrides lr/step (CUDA-graph replay) */
  int64_t kind; /* 0 adam (L2 wd folded in grad), 1 adamw (decoupled), 2 sgd nesterov (m = momentum buf, beta1 = momentum) */
  /* RESIDENT-PACKED mode (pack_idx != NULL): p, g, m, v are all in the packed K-major layout of the GEMM matrices (p = fp32
   * master, g = what the wgrad kernels accumulated), so every access is coalesced; positions with pack_idx[j] < 0 (structural
   * zeros / padding) keep their p, m, v.  The master is also written as the operand copy the next forward GEMMs load, at
   * EVERY position (padding holds zeros in the master): packed_out[j] (fp32, rounded to TF32 if flags & SCV_F_ROUND_TF32)
   * and / or packed16_out[j] (bf16).  n must be a multiple of 4 and the arrays 16-byte aligned (packed16_out: 8).  No
   * repack pass and no gradient-unpack pass is left in the step. */
  const int32_t* pack_idx; float* packed_out; void* packed16_out; int64_t flags;
  /* optional liveness bitmask of the packed positions (bit j of word w: position 32 w + j is live), ceil(n / 32) words:
   * when given it replaces the sign test on pack_idx (which may then be NULL) - 1/32 of the index array's traffic */
  const uint32_t* pack_mask;
} scv_optim_t;
int scv_optim_step(const scv_optim_t* p, void* stream);
/* global gradient norm over the packed gradients: sumsq[0] += sum over the live j < n (pack_mask bit, or pack_idx[j] >= 0 when
 * pack_mask is NULL) of gpacked[j]^2 */
int scv_sumsq_packed(const float* gpacked, const int32_t* pack_idx, const uint32_t* pack_mask, int64_t n, double* sumsq,
                     void* stream);
/* cudaMemsetAsync(p, 0, bytes) on `stream`: the per-step accumulators (BatchNorm sums, loss terms, grad norm) live in
 * one buffer and are cleared by one memset node of the step's CUDA graph. */
int scv_zero(void* p, int64_t bytes, void* stream);

/* out[i] = float(in[i]) */
int scv_d2f(const double* in, float* out, int64_t n, void* stream);

/* get_batch_loss total, train/losses.py:320-322: out[i] = float(acc[i]) for i < n and
 * out[n] = sum_i scale[i] * acc[i] over scale[i] != 0 (one thread; n <= 64). */
int scv_loss_finalize(const double* acc, const float* scale, float* out, int64_t n, void* stream);

/* ResVAE.decode model/residual.py:484-489: root_hat[f][d] = inv_normalize_root(xh[f][nx + d]) */
int scv_unpack_root(const float* xh, int64_t ld, int64_t nx, const float* arena, float* root_hat, int64_t F,
                    void* stream);

/* ---- pose-window preprocessing: preprocess_save_data data/dataset.py:313-454 ----------------------
 * The raw keypoint frames `pose` (N x J x 3 float64) stay resident in HBM; a window is W consecutive frames
 * starting at starts[w] (run boundaries at animal-id changes are host metadata, get_window_indices :198-233).
 * scv_window_indices : winds[w][t] = starts[w] + t (int64; bit-exact with the reference index matrix)
 * scv_window_features: per window, in float64 like numpy: speed[w] = mean keypoint speed (get_speed_outliers
 *   :299-309; the caller drops windows with speed > threshold), avg_speed_3d[w] = [root, parts[0], mean of the
 *   other parts] (get_speed_parts :134-163 + :362-374; parts = [n, len, j.., len, j..], the first joint of each part
 *   is its reference), yaw[w] of the mid frame's joint0->joint1 direction (:236-243), heading = (sin, cos) (:260-267)
 * scv_preprocess_windows: for the kept windows keep[i] (NULL = all): root centring + mid-forward rotation
 *   (:383-413), inv_kin (:11-46) -> local quaternions -> 6-D (quaternion.py:325-334), integer-truncated segment
 *   offsets (:279-296), target_pose = FK with zero root (:438-449).  mode 0 none, 1 midfwd, 2 x360 (centre only).
 *   tree = [n_chains, len, j.., ...]; offset = J x 3 int32 unit offsets (configs/mouse_skeleton.yaml:95-112).
 *   Outputs are (n_keep, W, J, 6|3) / (n_keep, W, 3) float32, contiguous. */
int scv_window_indices(const int64_t* starts, int64_t n_w, int64_t window, int64_t* winds, void* stream);
int scv_window_features(const double* pose, const int64_t* starts, int64_t n_w, int64_t window, int64_t J,
                        const int32_t* parts, double* speed, float* avg_speed_3d, float* heading, double* yaw,
                        void* stream);
int scv_preprocess_windows(const double* pose, const int64_t* starts, const int64_t* keep, int64_t n_keep,
                           int64_t window, int64_t J, const int32_t* tree, const int32_t* offset, const double* yaw,
                           int64_t mode, float* x6d, float* root, float* offsets, float* target_pose, void* stream);

/* ---- "mcmi" scrubbing loss: kernel mutual-information estimate between the latent mean and the conditioning variables
 * (MutInfoEstimator model/disentangle.py:234-317, loss train/losses.py:221-225, estimator rebuild after every optimizer
 * step train/trainer.py:184-199).  Stored samples xs (S,z), ys (S,dy); var_s (S,z) for var_mode "diagonal" (NULL = "sphere":
 * the scalar bandwidth); logAx (S) (diagonal) or (1) (sphere); valid: device flag, 0 = no estimator yet (loss 0, no gradient).
 * scv_mi_loss: loss[0] += mean_b [LSE_s a_xy - LSE_s a_x - LSE_s a_y] (double, if loss != NULL); dx (B,z) += gscale[0] * d loss/d x
 * (if dx != NULL; gscale NULL = 1).  y rows have stride y_ld.
 * scv_mi_update: xs = mu, ys = var rows, var_s = diag(L)^2 + bandwidth, logAx, valid = 1. */
int scv_mi_loss(const float* x, const float* y, int64_t y_ld, const float* xs, const float* ys, const float* var_s,
                const float* logAx, double bandwidth, int64_t S, int64_t B, int64_t z, int64_t dy, const float* valid,
                double* loss, const float* gscale, float* dx, void* stream);
int scv_mi_update(const float* mu, const float* L, const float* var, int64_t var_ld, float* xs, float* ys, float* var_s,
                  float* logAx, double bandwidth, int64_t S, int64_t z, int64_t dy, float* valid, void* stream);

/* ---- "moving_avg" scrubber (MovingAverageFilter model/disentangle.py:9-88; loss train/losses.py:286-289; update
 * train/trainer.py:169-178): per class c two running means m1, m2 (nc,z) of the latent mean with forgetting factors lam1 < lam2.
 * scv_ma_loss (evaluate_loss :32-74): class means xbar_c of x over the members of class c (stat: scratch nc (z + 1) floats, the
 *   member count last; an empty class gives NaN as torch.mean of an empty selection); ||xbar_c - m1_c|| < ||xbar_c - m2_c|| ?
 *   (lam1 = clamp(lam1 - delta), lam2 = lam1 + lamdiff) : (lam2 = clamp(lam2 + delta), lam1 = lam2 - lamdiff); estimates
 *   e_c = ((1 - lam1) xbar + lam1 m1 + (1 - lam2) xbar + lam2 m2) / 2; loss[0] += sqrt(sum_{c < c'} |e_c - e_c'|^2) (NOT divided by
 *   the batch size, as the reference); coef (nc,z) = d loss / d x_b for a member b of class c.
 * scv_ma_backward: dx (B rows of d_ld) += gscale[0] * coef[class of y_b].
 * scv_ma_update (:76-88): m_i = (1 - lam_i) xbar + lam_i m_i. */
int scv_ma_loss(const float* x, int64_t x_ld, const int64_t* y, const int64_t* classes, int64_t nc, int64_t z, int64_t B,
                const float* m1, const float* m2, float* lam1, float* lam2, double delta, double lamdiff, float* stat, float* coef,
                double* loss, void* stream);
int scv_ma_backward(const int64_t* y, const int64_t* classes, const float* coef, const float* gscale, int64_t nc, int64_t z,
                    int64_t B, float* dx, int64_t d_ld, void* stream);
int scv_ma_update(const float* x, int64_t x_ld, const int64_t* y, const int64_t* classes, int64_t nc, int64_t z, int64_t B,
                  const float* lam1, const float* lam2, float* m1, float* m2, float* stat, void* stream);

/* ---- "moving_avg_lsq" scrubber (MovingAvgLeastSquares model/disentangle.py:393-538, polynomial order 1; loss
 * train/losses.py:237-245; running-covariance update after the optimizer step train/trainer.py:169-178).
 * x = mu rows (B,z) with an appended column of ones when bias != 0: nx = z + bias <= 136; y (B,ny), ny <= 16.
 * scv_mals_solve: W_i = solve(Sxx_i + diag(l2_reg; the bias row excluded), Sxy_i), i = 0,1 (fp32, partial pivoting), W_i (nx,ny).
 * scv_mals_loss: l01[0] += sum (y - x W0)^2, l01[1] += sum (y - x W1)^2 (double; if l01 != NULL); yhat0/yhat1 (B,ny) optional
 *   outputs; dmu (B rows of d_ld) += gscale[0] / B * ((yhat0 - y) W0^T + (yhat1 - y) W1^T) over the first z rows (if dmu != NULL).
 * scv_mals_finalize (evaluate_loss :505-538): the forgetting factors move by delta towards the better decoder
 *   (l0 < l1: lam0 = clamp(lam0 - delta, 0, 1), lam1 = lam0 + lamdiff; else lam1 = clamp(lam1 + delta, 0, 1), lam0 = lam1 -
 *   lamdiff) and loss[0] += (l0 + l1) / 2 / B (if loss != NULL).
 * scv_mals_update (:489-503): Sxx_i = lam_i Sxx_i + x^T x, Sxy_i = lam_i Sxy_i + x^T y. */
int scv_mals_solve(const float* Sxx0, const float* Sxy0, const float* Sxx1, const float* Sxy1, double l2_reg, int64_t bias,
                   int64_t nx, int64_t ny, float* W0, float* W1, void* stream);
int scv_mals_loss(const float* mu, int64_t mu_ld, const float* y, int64_t y_ld, const float* W0, const float* W1, int64_t bias,
                  int64_t B, int64_t z, int64_t ny, double* l01, float* yhat0, float* yhat1, const float* gscale, float* dmu,
                  int64_t d_ld, void* stream);
int scv_mals_finalize(const double* l01, float* lam0, float* lam1, double delta, double lamdiff, int64_t B, double* loss,
                      void* stream);
int scv_mals_update(const float* mu, int64_t mu_ld, const float* y, int64_t y_ld, int64_t bias, int64_t B, int64_t z, int64_t ny,
                    const float* lam0, const float* lam1, float* Sxx0, float* Sxy0, float* Sxx1, float* Sxy1, void* stream);

/* ---- "qda" scrubber (QuadraticDiscriminantFilter model/disentangle.py:90-232; loss train/losses.py:247-251; update
 * train/trainer.py:169-178).  Per class c (nc <= 16 labels in `classes`, int64) two one-vs-rest Gaussian classifiers A, B, each
 * with (mean, covariance) of "label != c" (0) and "label == c" (1): m0a, m1a, m0b, m1b (nc,z), S0a, S1a, S0b, S1b (nc,z,z), z <= 128.
 * scv_qda_factor: SinvT (4,nc,z,z) = transposed inverses (Gauss-Jordan, partial pivoting, fp32), logdet (4,nc) (NaN if det < 0).
 * scv_qda_loss: x (B rows of x_ld), labels y (B, int64).  acc (double, 4 per class) += [lla, llb, llra, llrb] with
 *   cgll_q(x) = -1/2 (logdet_q + (x - m_q)^T S_q^-1 (x - m_q)), lla = sum_b cgll_{[y_b == c] a}, llra = sum_b s_b (cgll_1a - cgll_0a),
 *   s_b = +-1 (if acc != NULL); dx (B rows of d_ld) += gscale[0] / (2 nc B) sum_c s_b [(t_0a - t_1a) + (t_0b - t_1b)],
 *   t_q = S_q^-1 (x_b - m_q) (if dx != NULL).
 * scv_qda_finalize (evaluate_loss :173-232): per class lla > llb ? (lama = clamp(lama - delta), lamb = lama + lamdiff)
 *   : (lamb = clamp(lamb + delta), lama = lamb - lamdiff); loss[0] += sum_c (llra + llrb) / 2 / nc / B (if loss != NULL).
 * scv_qda_update (:133-171): class-conditional batch means and covariances (correction 0; an empty subset gives NaN, as
 *   torch.mean of an empty selection) blended into the running buffers with lama (A) and lamb (B); stat: scratch of
 *   2 nc (z + 1) floats. */
int scv_qda_factor(const float* S0a, const float* S1a, const float* S0b, const float* S1b, int64_t nc, int64_t z, float* SinvT,
                   float* logdet, void* stream);
int scv_qda_loss(const float* x, int64_t x_ld, const int64_t* y, const int64_t* classes, const float* m0a, const float* m1a,
                 const float* m0b, const float* m1b, const float* SinvT, const float* logdet, int64_t nc, int64_t z, int64_t B,
                 double* acc, const float* gscale, float* dx, int64_t d_ld, void* stream);
int scv_qda_finalize(const double* acc, float* lama, float* lamb, double delta, double lamdiff, int64_t nc, int64_t B, double* loss,
                     void* stream);
int scv_qda_update(const float* x, int64_t x_ld, const int64_t* y, const int64_t* classes, int64_t nc, int64_t z, int64_t B,
                   const float* lama, const float* lamb, float* m0a, float* m1a, float* m0b, float* m1b, float* S0a, float* S1a,
                   float* S0b, float* S1b, float* stat, void* stream);

/* ---- generative restrictiveness (eval): reference eval/eval.py:22-120.  For B decoded windows (xh rows of ld floats: the
 * decoder output after tanh, 6-D channels first; root_hat (B*W,3) un-normalised root positions or NULL = 0; offsets
 * (B*W,J,3)): forward kinematics (fwd_kin_cont6d_torch data/dataset.py:83-116, eps 1e-8) and, per window,
 *   heading (B,2)  = (sin, cos) of yaw = -atan2 of the unit joint0->joint1 vector at the mid frame (:67-72)
 *   avg3 (B,3)     = [mean root speed, mean speed of part 0's joints, mean of the means of parts 1 and 2] (:73-104),
 *                    then (x - norm[0..3)) / norm[3..6) if norm != NULL (:105-118)
 * parts = [n, len, ref, j.., len, ref, j..]; pose_out (B,W,J,3) optional.  Each may be NULL. */
int scv_gen_features(const float* xh, int64_t ld, const float* root_hat, const float* offsets, const int32_t* tree,
                     int64_t n_tree, const int32_t* parts, int64_t B, int64_t W, int64_t J, const float* norm,
                     float* pose_out, float* heading, float* avg3, void* stream);

#ifdef __cplusplus
}
#endif
#endif
