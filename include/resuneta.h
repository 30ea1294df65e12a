/* resuneta.h — C-ABI of libresuneta.so: the B200 (sm_100a) kernels behind the ResUnet-a
 * multitask hot path.  Plain pointers and sizes only; all tensors are caller-owned device
 * memory, NHWC, dense.  Every launch is asynchronous on the given cudaStream_t (passed as
 * void*), never synchronises the host and never allocates.  Return value: 0 on success,
 * <0 = RSA_ERR_*; rsa_last_error() returns a thread-local message.
 *
 * Each entry point replaces work that the reference delegates to TensorFlow/Keras library
 * kernels (the reference has no native code of its own); the call sites it
 * stands in for are cited per function as  file:line  relative to the reference repo.
 *
 * dtype codes: RSA_F32 = 0 (validation mode), RSA_BF16 = 1 (performance mode).
 */
#ifndef RESUNETA_H
#define RESUNETA_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define RSA_F32 0
#define RSA_BF16 1

#define RSA_OK 0
#define RSA_ERR_SHAPE (-1)
#define RSA_ERR_DTYPE (-2)
#define RSA_ERR_ALIGN (-3)
#define RSA_ERR_ARCH (-4)
#define RSA_ERR_CUDA (-5)

#define RSA_MAX_SEG 16

/* One K-segment of an implicit GEMM: `C` channels gathered from `src` (N,Hs,Ws,C) at
 *   hs = ((h*mult) >> shift) + off_h ,  ws = ((w*mult) >> shift) + off_w
 * for output pixel (n,h,w); out-of-range source pixels contribute 0 ('same' zero padding is
 * applied AFTER any preceding BN+ReLU, model2.py:17-20).  aligned!=0 additionally requires
 * (h*mult) and (w*mult) to be multiples of 1<<shift (transposed stride-2 gather).
 * relu_in!=0 applies max(.,0) to the gathered value (combine, model2.py:82).
 * w_off = element offset of this segment's weight block inside `w`. */
typedef struct {
  const void* src;
  int32_t C, Hs, Ws;
  int32_t mult, shift, off_h, off_w;
  int32_t relu_in, aligned;
  int64_t w_off;
} rsa_seg_t;

const char* rsa_version(void);
const char* rsa_last_error(void);
/* 0 if the current device is sm_100 (B200); RSA_ERR_ARCH otherwise. */
int rsa_device_check(void);

/* ---- generic implicit GEMM (fp32 accumulate, CUDA cores) --------------------------------
 * out[m, co] = epi( sum_seg sum_c gather(seg, m)[c] * B_seg[c, co] + bias[co] )
 *   transB==0: B_seg[c,co] = w[w_off + c*ldw + co]      (HWIO kernels / [K,Co] matrices)
 *   transB!=0: B_seg[c,co] = w[w_off + co*ldw + c]      (data gradients)
 * epi: += residual[m,co]; += old out (accumulate); relu; *= (mask[m,co] > 0); optional
 * per-channel sum / sum-of-squares of the stored values atomically added into stats[2*Co]
 * (double) — the statistics the next BatchNormalization needs.
 * Replaces: Conv2D 3x3 dilated 'same' (model2.py:19-24), Conv2D 1x1 / stride 2
 * (model2.py:37,84,92,101-111), UpSampling2D+Concatenate gathers (model2.py:55-76,83,91),
 * head convs (model2.py:153-188) and all their data gradients. */
int rsa_igemm_fwd(const rsa_seg_t* segs, int nseg, int in_dtype, const float* w, int ldw, int transB,
                  const float* bias, void* out, int out_dtype, const void* residual, const void* mask,
                  double* stats, int N, int Ho, int Wo, int Co, int accumulate, int relu, void* stream);

/* dw[w_off + c*ldw + co] += sum_m gather(seg,m)[c] * dy[m,co];  dbias[co] += sum_m dy[m,co]
 * (dbias may be NULL).  dw/dbias are fp32 and must be zeroed by the caller once per step. */
int rsa_igemm_wgrad(const rsa_seg_t* segs, int nseg, int in_dtype, const void* dy, int dy_dtype,
                    float* dw, int ldw, float* dbias, int N, int Ho, int Wo, int Co, void* stream);

/* ---- BatchNormalization (keras defaults eps=1e-3, momentum=.99; model2.py:17,21,38,86,93) --
 * Statistics travel as double[2*C] = {sum_c, sumsq_c} over `count` elements per channel.
 * stats==NULL selects inference mode (moving statistics). */
int rsa_bn_stats(const void* x, int dtype, int64_t M, int C, double* stats, void* stream);
/* nout outputs y_k = act(gamma_k * xhat + beta_k), all sharing the statistics of x
 * (the ResBlock-a branches normalise the same input, model2.py:17).  meaninv_out (optional, fp32 [2][C]) receives
 * {mean, invstd} of x: the table the fused BatchNorm-backward epilogue of rsa_conv_tc2_fwd reads. */
int rsa_bn_apply(const void* x, int dtype, int64_t M, int C, int nout, void* const* outs,
                 const float* const* gammas, const float* const* betas, const double* stats, double count,
                 const float* const* moving_means, const float* const* moving_vars, float eps, int relu,
                 float* meaninv_out, void* stream);
/* red[2*C] (double, zeroed by caller) += { sum g, sum g*xhat },  g = dy * (act>0 if act) */
int rsa_bn_bwd_reduce(const void* dy, const void* x, const void* act, int dtype, int64_t M, int C,
                      const double* stats, double count, float eps, double* red, void* stream);
/* dx (=|+=) gamma*invstd*(g - red0/count - xhat*red1/count); dgamma = red1, dbeta = red0 */
int rsa_bn_bwd_apply(const void* dy, const void* x, const void* act, int dtype, int64_t M, int C,
                     const double* stats, double count, float eps, const float* gamma, const double* red,
                     void* dx, int accumulate, float* dgamma, float* dbeta, void* stream);
/* Backward of k <= 4 BatchNormalization(+ReLU) layers that share their input x — the ResBlock-a branches all
 * normalise the same tensor (model2.py:17-18).  The ReLU mask is recomputed from x (bn_b(x) > 0).
 * reduce: reds[b][2C] (double, zeroed) += { sum g_b, sum g_b*xhat };
 * apply : dx (=|+=) sum_b gamma_b*invstd*(g_b - reds[b][0]/count - xhat*reds[b][1]/count), dgamma_b, dbeta_b. */
int rsa_bn_bwd_reduce_multi(const void* const* dys, const void* x, int dtype, int64_t M, int C, int k,
                            const double* stats, double count, float eps, const float* const* gammas,
                            const float* const* betas, int relu, double* const* reds, void* stream);
int rsa_bn_bwd_apply_multi(const void* const* dys, const void* x, int dtype, int64_t M, int C, int k,
                           const double* stats, double count, float eps, const float* const* gammas,
                           const float* const* betas, int relu, double* const* reds, void* dx, int accumulate,
                           float* const* dgammas, float* const* dbetas, void* stream);
/* out[2C] = {mean, invstd} in fp32 from the {sum, sumsq} statistics. */
int rsa_bn_meaninv(const double* stats, double count, float eps, float* out, int C, void* stream);
/* statistics of y = gamma*xhat+beta derived analytically from the statistics of x */
int rsa_bn_derive_stats(const double* src_stats, double count, const float* gamma, const float* beta,
                        float eps, double* dst_stats, double dst_count, int C, void* stream);
/* moving <- moving*momentum + batch*(1-momentum) for a table of BN layers in one launch.
 * table[i] = {stats_off, C, mean_off, var_off} (int64 x4); counts[i] = {count, count_full}
 * (count_full = element count of the tensor keras normalised; differs when the BN was evaluated
 * at pooled resolution).  TF feeds the Bessel-corrected variance. */
int rsa_bn_update_moving(const double* stats_base, float* param_base, const int64_t* table,
                         const double* counts, int nlayers, float momentum, void* stream);

/* ---- PSPPooling pyramid (model2.py:41-79) ---------------------------------------------------
 * One pass over x (N,H,W,C) producing the stride-k max-pools k=2,4,8 (NULL to skip a level). */
int rsa_maxpool_pyr_fwd(const void* x, int dtype, int N, int H, int W, int C, void* p2, void* p4, void* p8,
                        void* stream);
/* dx (=|+=) sum_k dp_k routed to the first maximum of each window (row-major scan order). */
int rsa_maxpool_pyr_bwd(const void* x, int dtype, int N, int H, int W, int C, const void* dp2,
                        const void* dp4, const void* dp8, void* dx, int accumulate, void* stream);
/* stride-k window sums k=2,4,8 (adjoint of nearest up-sampling, UpSampling2D model2.py:55-60) */
int rsa_sumpool_pyr(const void* x, int dtype, int N, int H, int W, int C, void* s2, void* s4, void* s8,
                    void* stream);

/* ---- head activations (model2.py:162,171,182,186) — fp32 [M,C], in place allowed ---------- */
int rsa_softmax_fwd(const float* z, float* p, int64_t M, int C, void* stream);
int rsa_softmax_bwd(const float* p, const float* dp, float* dz, int64_t M, int C, void* stream);
int rsa_sigmoid_fwd(const float* z, float* p, int64_t n, void* stream);
int rsa_sigmoid_bwd(const float* p, const float* dp, float* dz, int64_t n, void* stream);

/* ---- Tanimoto dual loss (multitasking_utils.py:38-85) ---------------------------------------
 * sums[B,C,5] (double, zeroed) += { Σp, Σp², Σl, Σl², Σp·l } over H*W */
int rsa_tanimoto_sums(const float* pred, const float* label, int B, int64_t HW, int C, double* sums,
                      void* stream);
/* loss_b[B]; loss_mean[1] = mean_b loss_b; coef[B,C,3] with
 * d(scale*loss_mean)/d pred[b,x,c] = coef0 + coef1*pred + coef2*label (gradient flows through the
 * prediction-derived class weights of the first term, multitasking_utils.py:79). */
int rsa_tanimoto_finalize(const double* sums, int B, int64_t HW, int C, float scale, float* loss_b,
                          float* loss_mean, float* coef, void* stream);
int rsa_tanimoto_bwd(const float* pred, const float* label, const float* coef, int B, int64_t HW, int C,
                     float* dpred, void* stream);

/* ---- element-wise losses, keras SUM_OVER_BATCH_SIZE mean over [B,H,W] -------------------------
 * kind: 0 weighted/plain categorical CE (utils.py:466-491; weights NULL = 1), 1 binary CE,
 * 2 mean squared error (train_ISPRS.py:426-428).  loss_sum[1] (double, zeroed) += Σ_pixels loss. */
int rsa_pixel_loss_fwd(int kind, const float* pred, const float* label, const float* weights, int64_t M,
                       int C, double* loss_sum, void* stream);
int rsa_pixel_loss_bwd(int kind, const float* pred, const float* label, const float* weights, int64_t M,
                       int C, float scale, float* dpred, void* stream);
/* out[M] = un-reduced per-pixel loss: what the loss callables return when called directly,
 * weighted_categorical_crossentropy(w)(y_true, y_pred) -> [B,H,W] (utils.py:478-491). */
int rsa_pixel_loss_elem(int kind, const float* pred, const float* label, const float* weights, int64_t M,
                        int C, float* out, void* stream);

/* seg metrics (train_ISPRS.py:446-449): out[5] (int64, zeroed) += {argmax matches, TP, FP, TN, FN}@0.5 */
int rsa_seg_metrics(const float* pred, const float* label, int64_t M, int C, int64_t* out, void* stream);

/* ---- optimizers (train_ISPRS.py:404-407; keras forms) ---------------------------- */
/* lr_dev != NULL: the step size is read from device memory (lets a captured CUDA graph be replayed
 * while keras' bias-corrected lr_t / a changed optimizer.lr varies from step to step). */
int rsa_adam_step(float* param, const float* grad, float* m, float* v, int64_t n, float lr_t,
                  const float* lr_dev, float beta1, float beta2, float eps, float grad_scale, void* stream);
int rsa_sgd_step(float* param, const float* grad, float* vel, int64_t n, float lr, const float* lr_dev,
                 float momentum, float grad_scale, void* stream);

/* ---- inference side (test_ISPRS.py:295-314) -----------------------------------------------------
 * pred_label[m] = argmax_c prob[m,c] (first maximum, numpy semantics); optional confusion matrix
 * cm[K,K] (int64, zeroed; rows = true label) += counts, with true_label int32 in [0,K). */
int rsa_argmax_confusion(const float* prob, int64_t M, int C, int32_t* pred_label, const int32_t* true_label,
                         int K, int64_t* cm, void* stream);

/* ---- tensor-core (tcgen05 + TMA) convolutions, bf16 mode --------------------------------------------------
 * 1 if the 3x3 layer (Cin, Cout at N x H x W) is taken by the tensor-core kernels (rsa_conv_tc2_fwd / rsa_conv_tc3_fwd
 * forward and data gradient, rsa_conv_tc_wgrad / rsa_conv_tc3_wgrad weight gradient) that replace cuDNN's dilated Conv2D
 * behind keras at model2.py:19-24,153-178. */
int rsa_conv_tc_supported(int N, int H, int W, int Cin, int Cout);
/* Persistent tcgen05 / TMA implicit-GEMM convolution (conv_tc2.cu): the 3x3 layers with C >= 128 and every 1x1 layer:
 * out[n,h,w,:Cout] = epi( sum_tap sum_src sum_c x_src[n, h*s+dy*dil, w*s+dx*dil, c] * wt[tap][co][koff_src+c]
 *                        + bias + sum_u up_{shift_u}(q_u) )
 * x0 / optional x1: bf16 NHWC sources, K-concatenated (taps = 1; Concatenate model2.py:83) or the single 3x3 source
 * (taps = 9); wt bf16 [taps][CoutP][C0+C1], CoutP = Cout rounded up to the N tile with zero rows; in_stride 2 =
 * Conv2D strides=2 (model2.py:103-111); q_u bf16 [N, H>>shift, W>>shift, Cout] are low-resolution addends that are
 * nearest-up-sampled in the epilogue (UpSampling2D, model2.py:55-60,91); out bf16 or fp32.  The sources use columns
 * [k_base, k_base+C0+C1) of wt's K dimension of length k_total (0 = C0+C1); out_stride 2 scatters the result to the even
 * pixels of a (2H,2W) tensor (data gradient of the stride-2 convolutions).  bnr_x / bnr_coef select the fused
 * BatchNormalization-backward reduction epilogue: `stats` += {sum g, sum g*xhat}, xhat = (bnr_x - mean)*invstd with
 * bnr_coef = rsa_bn_meaninv()'s [2][Cout] table (FusedBatchNormGrad's reductions, model2.py:17,21).
 * taps = 1 launches with at most 64 input channels in total and 8 / 16 / 32 / 64 bf16 output channels (the thin 256x256 /
 * 128x128 layers, HBM-bound) are executed by the streaming kernel of pw_stream.cu - same arguments, same epilogue order. */
int rsa_conv_tc2_supported(int N, int H, int W, int C0, int C1, int Cout);
int rsa_conv_tc2_fwd(const void* x0, int C0, const void* x1, int C1, const void* wt, int CoutP, const float* bias,
                     void* out, int out_f32, const void* residual, const void* mask, double* stats, int N, int H,
                     int W, int Cout, int taps, int dil, int in_stride, int nup, const void* const* up_ptrs,
                     const int* up_shifts, int k_base, int k_total, int out_stride, const void* bnr_x,
                     const float* bnr_coef, int accumulate, int relu, void* stream);
/* Thin-layer (32- / 64-channel) dilated 3x3 'same' convolution (conv_tc3.cu): resident weights, halo tiles for |dil| <= 3,
 * up to four branches accumulated in one TMEM tile (ResBlock-a branch sum, model2.py:23-31), TMA-stored output and
 * BatchNormalization statistics computed by the tensor core.
 *   out = epi( sum_b conv3x3(xs[b], wts[b], dils[b]) + sum_b biases[b] ), epi: + residual, + out (accumulate), ReLU,
 *   mask (keep where mask > 0);  xs[b] bf16 [N,H,W,32]; wts[b] bf16 [9][32][32] as rsa_conv_tc2_fwd (negative dilation
 *   with the [tap][ci][co] copy = data gradient); stats double[2C] += {sum, sumsq} of the stored values;
 *   bnr_x != NULL (data gradient into a = [relu](BatchNorm(bnr_x))): the BatchNormalization backward reductions are fused:
 *   out = d(a) * relu-mask recomputed from bnr_x, stats += {sum g, sum g*xhat} (residual/accumulate still allowed, no mask).
 * Replaces cuDNN's Conv2D forward / backward-data behind model2.py:19-24,153-178 for the C = 32 layers. */
int rsa_conv_tc3_supported(int N, int H, int W, int C);
int rsa_conv_tc3_fwd(const void* const* xs, const void* const* wts, const float* const* biases, const int* dils, int nbr,
                     void* out, const void* residual, const void* mask, double* stats, int N, int H, int W, int C,
                     int accumulate, int relu, const void* bnr_x, const double* bnr_stats, double bnr_count, float bnr_eps,
                     const float* bnr_gamma, const float* bnr_beta, int bnr_relu, void* stream);
/* Weight gradient of the thin layers: one x halo per 16x16 item, the three taps of a tap row are the M atoms of one
 * MMA (conv_tc3.cu); same contract as rsa_conv_tc_wgrad, Cin == Cout == C in {32, 64}, dil > 0 (C = 64: dil <= 3). */
int rsa_conv_tc3_wgrad_supported(int N, int H, int W, int C, int dil);
int rsa_conv_tc3_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int C, int dil, void* stream);
/* dw[tap][ci][co] (fp32 HWIO, zeroed by the caller once per step) += sum_pix x[pix+off(tap), ci] * dy[pix, co];
 * x, dy bf16 NHWC, Cin == Cout.  Replaces cuDNN's Conv2D backward-filter behind model2.py:19-24,153-178. */
int rsa_conv_tc_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout, int dil,
                      void* stream);
/* dw[k*ldw + co] (fp32, zeroed by the caller once per step) += sum_pix x[n, h*s, w*s, k] * dz[n,h,w,co]: weight gradient
 * of one source of a 1x1 convolution on the tensor cores; Cin, Cout powers of two >= 16, s = in_stride (1|2).
 * Replaces the Conv2D 1x1 backward-filter behind model2.py:37,84,92,101-111. */
int rsa_pw_wgrad_tc(const void* x, const void* dz, float* dw, int ldw, int N, int H, int W, int Cin, int Cout,
                    int in_stride, void* stream);
/* db_k[c] += sum_m dy[m,c] for up to four bias gradients (NULL to skip): all branches of a ResBlock-a share
 * d(out) (Add, model2.py:27-31). */
int rsa_bias_grad(const void* dy, int dtype, long long M, int C, float* db0, float* db1, float* db2, float* db3,
                  void* stream);
/* bf16 copies of the fp32 HWIO master kernels for the tensor-core path, all layers in one launch.
 * table (device): nlayers entries {int64 src_off, int64 fwd_off, int64 bwd_off, int32 taps, Cin, Cout, pad};
 * fwd copy is [tap][Cout][Cin], bwd copy is [tap][Cin][Cout].  nlayers <= 512; max_elems = elements of the largest layer
 * (bounds the grid; the 32x32 tiles of all layers are walked as one flat list by a single wave of blocks). */
int rsa_pack_weights_tc(const float* params, void* shadow, const void* table, int nlayers, long long max_elems,
                        void* stream);

/* Head forward (bf16 mode): z[m, 0:n] fp32 = h[m, 0:32] bf16 . w[32][n] + b, n <= 8; the final Conv2D 1x1 of the
 * heads, model2.py:159,168,180,186. */
int rsa_head_fwd(const void* h, const float* w, const float* b, float* z, int64_t M, int n, void* stream);
/* ---- thin 1x1 convolutions: one side <= 16 channels, the other exactly 32 (thin.cu) --------------------------------
 * Stem Conv2D(32,(1,1)) on the raw n-band input (model2.py:101): out[m,0:32] = x[m,0:n].w[n][32] + b, optional BatchNorm
 * statistics of the output (double[64]); its weight/bias gradient; and the backward of a head's final 1x1 convolution
 * (model2.py:159,168,180,186) from the fp32 logits gradient dz: dh (=|+=) mask(h>0)*dz.w^T, dw[32][n] += h^T dz, db += sum dz. */
int rsa_stem_fwd(const void* x, int x_dtype, const float* w, const float* b, void* out, int out_dtype, int64_t M, int n,
                 double* stats, void* stream);
int rsa_stem_wgrad(const void* x, const void* dy, int dtype, int64_t M, int n, float* dw, float* db, void* stream);
int rsa_head_bwd(const void* h, int h_dtype, const float* dz, const float* w, int64_t M, int n, void* dh, int accumulate,
                 int relu_mask, float* dw, float* db, void* stream);

/* dst (=|+=) src, identity branch of the ResBlock-a backward (Add, model2.py:31) */
int rsa_axpy(void* dst, const void* src, int dtype, int64_t n, int accumulate, void* stream);

/* dtype conversion helper (host tensors arrive as fp32, train_ISPRS.py:122-141) */
int rsa_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream);

/* ---- multitask label generation (labels.cu; reference: OpenCV calls in multitasking_utils.py:6-34 and
 * preprocess_save_patches_ISPRS.py:89-94,224-228).  label / bound / dist: fp32 [N,H,W,C]; workspace:
 * rsa_label_workspace_bytes(N,H,W,C) bytes of device memory; rgb: uint8 [npix,3], color: fp32 [npix,3]. */
int64_t rsa_label_workspace_bytes(int N, int H, int W, int C);
int rsa_label_boundary(const float* label, float* bound, void* workspace, int N, int H, int W, int C, void* stream);
int rsa_label_distance(const float* label, float* dist, void* workspace, int N, int H, int W, int C, void* stream);
int rsa_label_hsv(const uint8_t* rgb, float* color, int64_t npix, void* stream);

/* ---- Amazon deforestation evaluation post-processing (postproc.cu; utils.py:505-548, utils2.py:312-356) ---- */
/* out = skimage.morphology.area_opening(img, area_threshold, connectivity=1) of a binary uint8 [H,W] map; workspace 2*H*W int32 */
int rsa_area_opening_binary(const uint8_t* img, uint8_t* out, int H, int W, int area_threshold, void* workspace, void* stream);
/* mask pipeline + 3x3 confusion counts (int64[9], zeroed by the caller) over the pixels under consideration, utils.py:527-545 */
int rsa_amazon_consider(const uint8_t* pred, const uint8_t* opened, const uint8_t* ref_clip, const uint8_t* clip_mask,
                        uint8_t* ref_consider, uint8_t* pred_consider, uint8_t* selected, int64_t n, int64_t* cm, void* stream);

#ifdef __cplusplus
}
#endif
#endif
