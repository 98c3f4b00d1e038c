/* holocron_b200 — C ABI of the B200 (sm_100a) kernels behind the holocron.nn / holocron.ops / holocron.optim
 * hot path of frgfm/Holocron.
 *
 * The reference is pure Python/PyTorch and has no FFI of its own: every entry point below replaces
 * the chain of ATen/cuDNN kernels that a reference *Python* function launches; the reference file:line each one
 * stands in for is cited per declaration (paths relative to the reference repository root).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - every pointer is a DEVICE pointer unless stated otherwise; `stream` is a cudaStream_t;
 *   - return value: 0 on success, otherwise a cudaError_t (launch-configuration errors included);
 *   - re-entrant, no global mutable state apart from one-time kernel attribute setup;
 *   - dtype codes: 0 = float32, 1 = bfloat16, 2 = float16;
 *   - activation tensors of the convolution / BatchNorm entry points are NHWC bf16 ("channels_last"),
 *     filters are KRSC ([Cout][R][S][Cin]).
 */
#ifndef HOLOCRON_B200_H
#define HOLOCRON_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- activations: holocron/nn/functional.py:30-41 (hard_mish), :44-56 (nl_relu) ------------------------ */
int hb_hard_mish_fwd(const void* x, void* y, size_t n, int dtype, void* stream);
int hb_hard_mish_bwd(const void* x, const void* dy, void* dx, size_t n, int dtype, void* stream);
int hb_nl_relu_fwd(const void* x, void* y, size_t n, float beta, int dtype, void* stream);
int hb_nl_relu_bwd(const void* x, const void* dy, void* dx, size_t n, float beta, int dtype, void* stream);
/* backward of the in-place variant, from the OUTPUT y = log(1 + beta*relu(x)) */
int hb_nl_relu_bwd_from_out(const void* y, const void* dy, void* dx, size_t n, float beta, int dtype, void* stream);

/* ---- dense convolutions (tcgen05 implicit GEMM): nn.Conv2d call sites of holocron/models/utils.py:71-76
 *      (conv_sequence), models/classification/repvgg.py:55-73 (RepBlock) and their autograd backward -------- */
/* y[N,Ho,Wo,Cout] = act(conv(x[N,H,W,Cin], w[Cout,R,S,Cin]) + bias + residual); Cin % 8 == 0, Cout % 16 == 0.
 * bias: fp32 [Cout] or NULL; residual: bf16 like y or NULL; act: 0 none, 1 relu; num_ctas: 0 = one per SM. */
int hb_conv2d_fprop_bf16(const void* x, const void* w, void* y, const float* bias, const void* residual, int N, int H,
                         int W, int Cin, int Cout, int R, int S, int stride, int pad, int dil, int act, int num_ctas,
                         void* stream);
/* General form of the above (one kernel launch), used by the fused RepVGG block (repvgg.py:71-73) and by every
 * conv -> BatchNorm2d unit of conv_sequence (utils.py:71-76) in training mode:
 *   K extension   xe != NULL: y += conv1x1(xe [N,Ho,Wo,Ce], we [Cout,1,1,Ce]) in the SAME accumulator (stride-1 layers:
 *                 the input gradient of both RepVGG branches, dX = dgrad3x3(dY3) + dgrad1x1(dY1) (+ residual));
 *   dual output   w2 != NULL: y2 [N,Ho,Wo,Cout] = conv1x1(x, w2 [Cout,1,1,Cin]; same stride, pad 0) from the centre-tap
 *                 loads of the RxS convolution (pad must be (R/2)*dil): x is read once for both RepVGG branches;
 *   statistics    stats / stats2 != NULL: per-channel (sum, sum of squares) partials of the bf16 outputs y / y2 as
 *                 float [*stat_slots][Cout][2]; the caller allocates hb_conv_stat_slots_max() slots, the launch writes
 *                 the first *stat_slots (host int) completely; summing the slots in order is deterministic. This replaces
 *                 the separate statistics pass of training-mode BatchNorm2d.
 *   patch norm    norm_mean != NULL (NormConv2d, holocron/nn/functional.py:322-413): y = norm_rstd[m] * (acc - norm_mean[m] *
 *                 norm_wsum[co]) + bias, m = output pixel: the per-patch standardisation of the im2col rows folded
 *                 algebraically into the epilogue (statistics from hb_patch_stats_bf16).
 * bias / residual / act apply to y only. Fields not used must be zero. */
typedef struct hb_conv_args {
  const void* x; const void* w; void* y; const float* bias; const void* residual;
  int N, H, W, Cin, Cout, R, S, stride, pad, dil, act, num_ctas;
  const void* xe; const void* we; int Ce;
  const void* w2; void* y2;
  float* stats; float* stats2;
  const float* norm_mean; const float* norm_rstd; const float* norm_wsum;
} hb_conv_args;
int hb_conv2d_fused_bf16(const hb_conv_args* args, int* stat_slots, void* stream);
int hb_conv_stat_slots_max(void);
/* Per-patch statistics of the im2col rows of x [N,H,W,C] bf16 (zero padding included, like F.unfold): for every output
 * pixel m, mean[m] and rstd[m] = 1/sqrt(biased var + eps) over its K = k_logical = Cin*kh*kw patch values (channels
 * beyond the logical ones are zero padding of the layout and do not count). scratch: float [2*N*H*W]. */
int hb_patch_stats_bf16(const void* x, float* mean, float* rstd, float* scratch, int N, int H, int W, int C, int kh, int kw,
                        int stride, int pad, int dil, int k_logical, float eps, void* stream);
/* y = conv3x3(x, w; stride 1, pad 1) + sum_{e<nextra} conv1x1(xe_e, we_e), one accumulator (nextra <= 2; all inputs
 * [N,H,W,Cin] bf16, w [Cout,3,3,Cin], we_e [Cout,1,1,Cin]). Input gradient of a RepVGG block in one kernel
 * (holocron/models/classification/repvgg.py:71-73). Returns cudaErrorNotSupported (801) when the filter does not fit the
 * shared-memory-resident scheme (Cin, Cout <= 64 typically) - fall back to separate convolutions then. */
int hb_conv3x3_accum_bf16(const void* x, const void* w, const void* xe0, const void* we0, const void* xe1, const void* we1,
                          int nextra, void* y, int N, int H, int W, int Cin, int Cout, int num_ctas, void* stream);
/* dw[Cout,R,S,Cin] (fp32, overwritten) = sum over pixels of dy[N,Ho,Wo,Cout] x im2col(x[N,H,W,Cin]).
 * workspace: optional fp32 scratch of hb_conv2d_wgrad_workspace_bytes(...) bytes: the per-pixel-range partial sums are
 * then reduced in a fixed order (deterministic); with NULL they are accumulated with fp32 atomics. */
size_t hb_conv2d_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, int dil,
                                       int num_ctas);
int hb_conv2d_wgrad_bf16(const void* x, const void* dy, float* dw, float* workspace, size_t workspace_bytes, int N, int H,
                         int W, int Cin, int Cout, int R, int S, int stride, int pad, int dil, int num_ctas, void* stream);
/* Both weight gradients of a stride-1 RepVGG block in ONE pass over x (the 1x1 branch reads x through the centre-tap
 * window of the rows the 3x3 branch already holds in shared memory; holocron/models/classification/repvgg.py:55-73):
 * dw = [dW3 (Cout,3,3,Cin) | dW1 (Cout,Cin)] fp32, overwritten. hb_repvgg_wgrad_workspace_bytes == 0 / return code
 * cudaErrorNotSupported (801): shape outside the row-window scheme, use hb_conv2d_wgrad_bf16 per branch. */
size_t hb_repvgg_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int num_ctas);
int hb_repvgg_wgrad_bf16(const void* x, const void* dy3, const void* dy1, float* dw, float* workspace, size_t workspace_bytes,
                         int N, int H, int W, int Cin, int Cout, int num_ctas, void* stream);
/* Accumulating forms: the fixed-order reduction ADDS onto dw / dw3 / dw1 (the parameters' .grad storage, e.g. views of
 * the flat data-parallel gradient bucket) instead of overwriting - what autograd's AccumulateGrad does with one more
 * element-wise kernel per parameter and step. hb_conv2d_wgrad_acc_bf16 returns cudaErrorNotSupported (801), dw untouched,
 * for shapes that run as a single pixel range. */
int hb_conv2d_wgrad_acc_bf16(const void* x, const void* dy, float* dw, float* workspace, size_t workspace_bytes, int N, int H,
                             int W, int Cin, int Cout, int R, int S, int stride, int pad, int dil, int num_ctas,
                             void* stream);
int hb_repvgg_wgrad_acc_bf16(const void* x, const void* dy3, const void* dy1, float* dw3, float* dw1, float* workspace,
                             size_t workspace_bytes, int N, int H, int W, int Cin, int Cout, int num_ctas, void* stream);
/* fp32 KRSC master filter -> bf16 KRSC [CoutF][R][S][CinP] (zero-padded rows / channels) and, if wd != NULL, the
 * flipped + transposed bf16 filter [CinD][R][S][CoutP] used by the data-gradient pass */
int hb_pack_conv_weights(const float* w, void* wf, void* wd, int Cout, int Cin, int R, int S, int CinP, int CinD,
                         int CoutP, int CoutF, void* stream);
/* The same for a whole network in ONE launch. metas: device array of hb_pack_meta_bytes()-byte rows
 * {const float* w; bf16* wf; bf16* wd (NULL allowed); int32 Cout, Cin, R, S, CinP, CinD, CoutP, CoutF}; chunks: device array
 * of int32 pairs (row index, chunk index), hb_pack_chunk_elems() output elements (wf then wd) per chunk. */
int hb_pack_conv_weights_multi(const void* metas, const void* chunks, int num_chunks, void* stream);
int hb_pack_chunk_elems(void);
int hb_pack_meta_bytes(void);
/* Data gradient of a stride-2 3x3 pad-1 convolution (the backward of the stride-2 nn.Conv2d at the head of every
 * RepVGG / Darknet / ReXNet stage, holocron/models/classification/repvgg.py:55-73, utils.py:28-86) without zero insertion:
 * the 4 parity classes of dx [N,H,W,Cd] are 1/2/2/4-tap correlations over dy [N,Ho,Wo,C] written to their sub-grids.
 * wcls: class filters from hb_pack_dgrad_s2_weights. dy1/wd1 (may be NULL): output gradient and [Cd,1,1,C] filter of a
 * parallel 1x1 stride-2 branch, accumulated into class (0,0). Ho = (H-1)/2+1, Wo = (W-1)/2+1. */
int hb_conv2d_dgrad_s2_bf16(const void* dy, const void* wcls, const void* dy1, const void* wd1, void* dx, int N, int H, int W,
                            int Ho, int Wo, int C, int Cd, int num_ctas, void* stream);
/* fp32 KRSC master filter [Cout,3,3,Cin] -> the four bf16 class filters [CinD][1+a][1+b][CoutP], (a,b) = (0,0), (0,1),
 * (1,0), (1,1), stored back to back (9*CinD*CoutP elements) */
int hb_pack_dgrad_s2_weights(const float* w, void* out, int Cout, int Cin, int CinD, int CoutP, void* stream);
/* y[N,Ho,Wo,C] = zeros, y[n, sp*p, sp*q, :] = x[n,p,q,:]  (input of a stride-sp transposed convolution) */
int hb_zero_insert_bf16(const void* x, void* y, int N, int Hi, int Wi, int Ho, int Wo, int C, int sp, void* stream);
/* NCHW image (dtype code) -> NHWC bf16 with channels zero-padded to CP (CP % 8 == 0) */
int hb_nchw_to_nhwc_pad_bf16(const void* x, void* y, int N, int C, int H, int W, int CP, int dtype, void* stream);

/* explicit im2col for stems (Cin <= 4): x NCHW (dtype code) -> col [N*Ho*Wo, Kp] bf16 with k = (r*S + s)*C + c, zero
 * padded to Kp (multiple of 8); the stem then runs as a 1x1 convolution over Kp channels */
int hb_im2col_smallc_bf16(const void* x, void* col, int N, int C, int H, int W, int R, int S, int stride, int pad, int Kp,
                          int dtype, void* stream);

/* ---- BatchNorm2d + branch sum + activation, fused: BatchNorm2d/act emitted by conv_sequence
 *      (holocron/models/utils.py:73-78) and the branch sum of RepBlock.forward (repvgg.py:71-73) ------------- */
/* Training-mode statistics are carried as PARTIALS: float [slots][C][2] = per-channel (sum, sum of squares) of a
 * subset of the rows, written by the producer of the tensor (hb_conv2d_fused_bf16's stats/stats2, hb_bn_act_fwd_bf16's
 * out_stats) or by this stand-alone pass over u [M,C] bf16 (capacity hb_bn_stat_slots_max() slots, *slots = host int
 * out). hb_bn_finalize adds the slots of each branch in a fixed order in fp64: no floating-point atomics anywhere, two
 * runs give bit-identical statistics. */
int hb_bn_stats_partials_bf16(const void* u, int M, int C, float* parts, int* slots, void* stream);
int hb_bn_stat_slots_max(void);
/* parts, slots, gamma, beta, running_mean/var, num_batches_tracked: HOST arrays of B entries (device pointers / ints;
 * pointer entries other than parts may be NULL). Outputs fp32 [B][C]. Updates the running statistics with `momentum`
 * (unbiased variance) and increments the int64 num_batches_tracked counters, like nn.BatchNorm2d in training mode. */
int hb_bn_finalize(const float* const* parts, const int* slots, const float* const* gamma, const float* const* beta,
                   float* const* running_mean, float* const* running_var, long long* const* num_batches_tracked,
                   float* mean, float* rstd, float* scale, float* shift, int B, int C, int C_logical, int M, float eps,
                   float momentum, void* stream);
/* channels in [C_logical, C) are zero padding (the parameter arrays hold C_logical entries): scale = shift = 0 */
int hb_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                      float eps, int C, int C_logical, float* scale, float* shift, float* mean, float* rstd,
                      void* stream);
/* out = act(sum_b (scale_b * u_b + shift_b) + residual)  [res_after = 1: act(sum_b ...) + residual, the shortcut of
 * holocron/models/classification/resnet.py:75-87 (_ResBlock.forward) used by the Darknet ResBlocks]; act: 0 none 1 relu 2 relu6 3 silu 4 leaky(slope) 5 mish
 * 6 hard_mish, 7 funnel: out = max(sum_b(...), residual) (FReLU, holocron/nn/modules/activation.py:58-82) */
/* out_stats (optional): (sum, sum of squares) partials of the bf16 OUTPUT, float [*out_stat_slots][C][2] (capacity
 * hb_bn_stat_slots_max()): the statistics of the identity-branch BatchNorm of the next RepVGG block, for free. */
int hb_bn_act_fwd_bf16(const void* u0, const void* u1, const void* u2, int B, const float* scale, const float* shift,
                       const void* residual, void* out, int M, int C, int act, float slope, int res_after,
                       float* out_stats, int* out_stat_slots, void* stream);
/* backward of the above; scratch: double [hb_bn_bwd_scratch_doubles(M, C, B)] (uninitialised); du_b/dres/dgamma/dbeta may
 * be NULL; gamma_grad_acc / beta_grad_acc: optional HOST arrays of B device pointers (entries may be NULL) to the fp32
 * [C_logical] gradient buffers of the BatchNorm weight / bias, which dgamma_b / dbeta_b are ADDED to (deterministic
 * fixed-order reductions, no atomics). */
size_t hb_bn_bwd_scratch_doubles(int M, int C, int B);
int hb_bn_act_bwd_bf16(const void* dout, const void* u0, const void* u1, const void* u2, int B, const float* scale,
                       const float* shift, const float* mean, const float* rstd, const void* residual, double* scratch,
                       void* du0, void* du1, void* du2, void* dres, float* dgamma, float* dbeta,
                       float* const* gamma_grad_acc, float* const* beta_grad_acc, int C_logical, int M, int C, int act,
                       float slope, int train, int res_after, void* stream);

/* ---- depth-wise k x k convolution (NHWC bf16; weights fp32 [C,K,K]): FReLU's conv (activation.py:71-73) and the
 *      ReXNet depth-wise stage (holocron/models/classification/rexnet.py:112-125) ------------------------- */
int hb_dwconv_fwd_bf16(const void* x, const float* w, const float* bias, void* y, int N, int H, int W, int C, int K,
                       int stride, int pad, void* stream);
int hb_dwconv_bwd_data_bf16(const void* dy, const float* w, void* dx, int N, int H, int W, int C, int K, int stride,
                            int pad, void* stream);
/* dw fp32 [C,K,K], db fp32 [C] or NULL; scratch: double[hb_dwconv_wgrad_scratch_doubles(C, K)] (per-block partial sums,
 * folded in a fixed order: deterministic); K in {1,3,5,7} */
size_t hb_dwconv_wgrad_scratch_doubles(int C, int K);
int hb_dwconv_bwd_weight_bf16(const void* x, const void* dy, float* dw, float* db, double* scratch, int N, int H, int W,
                              int C, int K, int stride, int pad, void* stream);

/* ---- global average pooling: holocron/nn/modules/downsample.py:58-74 ----------------------------------- */
int hb_gap_fwd_bf16(const void* x, void* y, int N, int HW, int C, void* stream);
int hb_gap_bwd_bf16(const void* dy, void* dx, int N, int HW, int C, void* stream);

/* ---- squeeze-excite gate: SEBlock.forward `x * y` followed by the block's activation,
 *      holocron/models/classification/rexnet.py:63-66, 125-131 --------------------------------------------- */
/* out[n,p,c] = act(x[n,p,c] * gate[n,c]);  x/out [N,HW,C] bf16, gate fp32 [N,C]; act codes as hb_bn_act_fwd_bf16 (0-6) */
int hb_gate_act_fwd_bf16(const void* x, const float* gate, void* out, int N, int HW, int C, int act, float slope,
                         void* stream);
/* dz = dout * act'(x*gate); dx = dz * gate (bf16); dgate[n,c] = sum_p dz * x (fp32, overwritten, deterministic) */
int hb_gate_act_bwd_bf16(const void* dout, const void* x, const float* gate, void* dx, float* dgate, int N, int HW, int C,
                         int act, float slope, void* stream);

/* ---- box operators: holocron/ops/boxes.py:16-211 (+ torchvision.ops.boxes.box_iou, boxes.py:11) ---------- */
/* mode: 0 IoU, 1 GIoU, 2 DIoU penalty rho^2/c^2, 3 DIoU loss (== the reference's ciou_loss, boxes.py:208-209),
 * 4 aspect-ratio consistency. boxes fp32 [M,4]/[N,4] xyxy; out fp32 [M,N]. */
int hb_box_pairwise(const float* boxes1, const float* boxes2, float* out, int M, int N, int mode, void* stream);
/* sets *flag (device int) to 1 when a box has x2 < x1 or y2 < y1 (box_giou's AssertionError, boxes.py:56-57) */
int hb_box_degenerate(const float* boxes, int n, int* flag, void* stream);
/* g1 [M,4], g2 [N,4] (either may be NULL) = gradients of sum(gout * op) for modes 0-3 */
int hb_box_pairwise_bwd(const float* boxes1, const float* boxes2, const float* gout, float* g1, float* g2, int M, int N,
                        int mode, void* stream);

/* ---- NormConv2d / Add2d: holocron/nn/functional.py:322-462 (_xcorr2d, norm_conv2d, add2d) ---------------- */
/* x fp32 NCHW, w fp32 [Cout,Cin,KH,KW], out fp32 [N,Cout,Ho,Wo]; mode 0 multiply-accumulate, 1 adder (-L1);
 * normalize: standardise every im2col patch (biased var + eps); mean/rstd: fp32 [N*Ho*Wo] (written when normalize). */
int hb_xcorr2d_fwd(const float* x, const float* w, const float* bias, float* out, float* mean, float* rstd, int N, int Cin,
                   int H, int W, int Cout, int KH, int KW, int stride, int pad, int dil, int mode, int normalize,
                   float eps, void* stream);
int hb_xcorr2d_wgrad(const float* x, const float* w, const float* g, const float* mean, const float* rstd, float* dw,
                     int N, int Cin, int H, int W, int Cout, int KH, int KW, int stride, int pad, int dil, int mode,
                     int normalize, float eps, void* stream);
int hb_add2d_dgrad(const float* x, const float* w, const float* g, float* dx, int N, int Cin, int H, int W, int Cout,
                   int KH, int KW, int stride, int pad, int dil, void* stream);

/* ---- DropBlock: holocron/nn/functional.py:465-500, nn/modules/dropblock.py:14-41 ------------------------ */
/* mask[N,H,W] = 1 - maxpool_bs(noise <= gamma); *kept (device float) = sum(mask). block_size must be odd. */
int hb_dropblock_mask(const float* noise, float* mask, float* kept, int N, int H, int W, int block_size, float gamma,
                      void* stream);
/* out = x * mask * (N*H*W / *kept) (no rescale when *kept == 0); x: [N,C,H,W] logical, physical NHWC if channels_last */
int hb_dropblock_apply(const void* x, void* out, const float* mask, const float* kept, int N, int C, int H, int W,
                       int channels_last, int dtype, void* stream);

/* ---- losses: holocron/nn/functional.py:59-113 (focal_loss), :540-613 (poly_loss), :503-537 (dice_loss) --- */
/* logits x are [N, K, S] (S = prod of spatial dims); kind 0 focal / 1 poly-1; loss_pos: float[N*S];
 * partials: double[2 * hb_loss_max_partials()] scratch; fwd_out: float[3] = {sum, #valid, mean}. */
int hb_loss_max_partials(void);
int hb_cls_loss_hard_fwd(const void* x, const long long* target, const float* weight, float* loss_pos, double* partials,
                         float* fwd_out, int N, int K, int S, int ignore_index, int kind, float gamma, float eps,
                         int dtype, void* stream);
/* reduction: 0 none (gout[N*S]), 1 mean, 2 sum (gout[1]); dx like x */
int hb_cls_loss_hard_bwd(const void* x, const long long* target, const float* weight, const float* gout,
                         const float* fwd_out, void* dx, int N, int K, int S, int ignore_index, int kind, float gamma,
                         float eps, int reduction, int dtype, void* stream);
int hb_poly_soft_fwd(const void* x, const void* soft, const float* weight, float* loss_pos, double* partials,
                     float* fwd_out, int N, int K, int S, int ignore_index, float eps, int dtype, void* stream);
int hb_poly_soft_bwd(const void* x, const void* soft, const float* weight, const float* gout, void* dx, int N, int K,
                     int S, int ignore_index, float eps, int reduction, int dtype, void* stream);
/* scratch: double[hb_dice_scratch_doubles(K)] (per-block partial sums, folded in a fixed order: deterministic);
 * out: float[1]; coef: float[2K] (input of hb_dice_bwd) */
size_t hb_dice_scratch_doubles(int K);
int hb_dice_fwd(const void* x, const void* target, const float* weight, double* scratch, float* out, float* coef, int N,
                int K, long long S, float gamma, float eps, int dtype, void* stream);
int hb_dice_bwd(const void* target, const float* coef, const float* gout, void* dx, int N, int K, long long S, int dtype,
                void* stream);

/* ---- optimizers: holocron/optim/adabelief.py:121-167, lamb.py:79-137, tadam.py:160-212 ----------------- */
/* metas: device table of T records {p, g, m, v, vmax, aux, ext, numel} (8 x 8 bytes each, fp32 tensors);
 * chunks: device int2[num_chunks] = {tensor index, chunk index}, chunk = hb_optim_chunk_elems() elements. */
int hb_optim_chunk_elems(void);
/* ctl (may be NULL): device control block of a captured training step (see hb_train_ctl_* below): the kernels then read
 * lr (and beta1 when >= 0) from it and skip the whole update while its skip flag is set. */
int hb_adabelief_step(const void* metas, const void* chunks, int num_chunks, float lr, float beta1, float beta2,
                      float eps, float weight_decay, int amsgrad, int step, const int* step_dev, const void* ctl,
                      void* stream);
/* AdamP, holocron/optim/adamp.py:144-191 (the reference scripts' default optimizer, references/classification/train.py:340):
 * two launches per group (moments + per-tensor <p,g>, ||p||^2, ||g||^2, <p,pt>; then the projected update), 40 B/parameter.
 * scratch: double [4*T]. */
int hb_adamp_step(const void* metas, const void* chunks, int num_chunks, int T, float lr, float beta1, float beta2, float eps,
                  float weight_decay, int amsgrad, float delta, int step, const int* step_dev, const void* ctl,
                  double* scratch, void* stream);
int hb_lamb_step(const void* metas, const void* chunks, int num_chunks, int T, float lr, float beta1, float beta2,
                 float eps, float weight_decay, float clip_lo, float clip_hi, double* scratch, void* stream);
int hb_tadam_step(const void* metas, const void* chunks, int num_chunks, int T, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int amsgrad, float dof, int step, const int* step_dev, double* scratch,
                  void* stream);
int hb_step_increment(int* step_dev, const void* ctl, void* stream);
/* The remaining optimizers of holocron/optim, same tables:
 * Adan, adan.py:145-199 - aux = prev_grad (read, never written: reference quirk), ext = exp_avg_delta, vmax = its running max;
 *   one launch, 40 B/parameter. */
int hb_adan_step(const void* metas, const void* chunks, int num_chunks, float lr, float beta1, float beta2, float beta3,
                 float eps, float weight_decay, int amsgrad, int step, const int* step_dev, const void* ctl, void* stream);
/* AdEMAMix, ademamix.py:138-176 - ext = exp_avg_slow; one launch, 36 B/parameter. */
int hb_ademamix_step(const void* metas, const void* chunks, int num_chunks, float lr, float beta1, float beta2, float beta3,
                     float alpha, float eps, float weight_decay, int step, const int* step_dev, const void* ctl,
                     void* stream);
/* LARS, lars.py:91-135 - m = momentum_buffer (NULL without momentum); first != 0 when this step creates the buffers; with
 *   weight decay the gradients are overwritten by g + wd * p like the reference's in-place add_. scratch: double [2*T]. */
int hb_lars_step(const void* metas, const void* chunks, int num_chunks, int T, float lr, float momentum, float dampening,
                 float weight_decay, int nesterov, int first, double* scratch, void* stream);
/* RaLars, ralars.py:56-140 - mode 0 rectified update (x r_t), 1 plain Adam ratio, 2 unadapted momentum (chosen on the host
 *   from the SMA length); aux = local_lr (1 element, written). scratch: double [2*T]. */
int hb_ralars_step(const void* metas, const void* chunks, int num_chunks, int T, float lr, float beta1, float beta2, float eps,
                   float weight_decay, float clip_lo, float clip_hi, int mode, float r_t, int step, double* scratch,
                   void* stream);
/* Lookahead.sync_params, wrapper.py:122-135 - p = fast weights, m = slow weights: slow += rate * (fast - slow); fast = slow */
int hb_lookahead_sync(const void* metas, const void* chunks, int num_chunks, float sync_rate, void* stream);

/* ---- training-loop control on the device: holocron/trainer/core.py:135-227 (_fit_epoch: NaN-loss skipping :153-159,
 *      per-iteration scheduler.step() :161; _backprop_step: gradient accumulation, clip_grad_norm_, optimizer.step :184-208)
 *      and :262-269 (OneCycleLR / CosineAnnealingLR) without host synchronisation, CUDA-graph replayable -------------- */
/* ctl: device block of hb_train_ctl_bytes() bytes, zero-initialised by the caller, then [0] = lr, [1] = -1:
 *   f32 lr | f32 beta1 (<0: keep) | i32 skip | i32 bad | i32 iter | i32 nan_run | i32 opt_steps | f32 grad_norm */
int hb_train_ctl_bytes(void);
/* after a micro-batch: remember a non-finite *loss (device fp32 scalar) when skip_nan != 0 */
int hb_train_ctl_observe(void* ctl, const float* loss, int skip_nan, void* stream);
/* phase 0, before the optimizer update: lr / beta1 = table[min(iter, n-1)] (table: n x {lr, beta1} fp32 or NULL), skip =
 * bad, nan_run counts consecutive skipped updates; phase 1, after it: opt_steps += !skip, the window's bad flag is cleared */
int hb_train_ctl_step(void* ctl, const float* table, int n, int phase, void* stream);
/* once per iteration (the reference's scheduler.step()): iter += 1 */
int hb_train_ctl_tick(void* ctl, void* stream);
/* torch.nn.utils.clip_grad_norm_(params, max_norm) on a flat fp32 gradient buffer (core.py:194,205): fixed-order global L2
 * norm + in-place scaling by min(1, max_norm / (norm + 1e-6)), two launches; scratch: double
 * [hb_grad_clip_partials_max()]; ctl (may be NULL) receives the norm. */
int hb_grad_clip_partials_max(void);
int hb_grad_clip_norm(float* grads, long long n, float max_norm, double* scratch, void* ctl, void* stream);

/* ---- bookkeeping (not part of the reference surface) ------------------------------------------------ */
long long hb_launch_count(void);      /* kernels launched through this library since the last reset */
void hb_launch_count_reset(void);
const char* hb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HOLOCRON_B200_H */
