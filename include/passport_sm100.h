/*
 * passport_sm100.h — C ABI of libpassport_sm100.so (B200 / sm_100a only).
 *
 * This is the drop-in boundary for the passport-layer hot path of kamwoh/DeepIPR.
 * The reference has no FFI of its own (it is pure PyTorch); each entry point below
 * replaces a group of ATen/cuDNN calls made by the reference's nn.Modules and names
 * the reference lines it stands in for.  The Python mirror of the reference's module
 * surface (deepipr_b200/layers.py) binds these symbols with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer owned by the
 *     caller unless stated otherwise; nothing is allocated or freed by the library;
 *   - activations are NHWC bf16 ("channels_last"); weights enter as fp32 OIHW (the
 *     layout of nn.Conv2d.weight) and are re-laid-out by pp_weight_prep();
 *   - per-channel vectors (gamma, beta, statistics, gradients of them) are fp32;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no entry
 *     point synchronises the device;
 *   - return value 0 = success, negative = PPStatus; pp_last_error() returns a
 *     thread-local message for the last non-zero return.
 *   - there is NO CPU path: every compute entry point fails with PP_ENODEVICE when no
 *     sm_100 device is present.
 */
#ifndef PASSPORT_SM100_H_
#define PASSPORT_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PP_ABI_VERSION 16

typedef enum PPStatus {
  PP_OK = 0,
  PP_EBADSHAPE = -1,    /* inconsistent / unsupported dimensions            */
  PP_EUNSUPPORTED = -2, /* valid request this build has no kernel for       */
  PP_ELAUNCH = -3,      /* CUDA launch or driver error (message has detail) */
  PP_EWORKSPACE = -4,   /* workspace pointer NULL or too small              */
  PP_ENODEVICE = -5,    /* no CUDA device / not sm_100                      */
  PP_EBADARG = -6       /* NULL where a pointer is required                 */
} PPStatus;

/* PP_NORM_GN: statistics per (sample, group) with `groups` groups of O/groups channels — nn.GroupNorm(o // 16, o)
 * (passportconv2d.py:59-60, conv2d.py:13-14) and, with groups == O, nn.InstanceNorm2d(o) (:61-62, conv2d.py:15-16).
 * Same arithmetic in train and eval mode (no running statistics). */
enum { PP_NORM_NONE = 0, PP_NORM_BN_TRAIN = 1, PP_NORM_BN_EVAL = 2, PP_NORM_GN = 3 };
enum { PP_ALGO_AUTO = 0, PP_ALGO_TCGEN05 = 1, PP_ALGO_SIMT = 2 };
/* PPConvDesc.flags, read by pp_conv_block_bwd: ADD the gradient into the caller's buffer instead of overwriting it
 * (what autograd's AccumulateGrad does with `param.grad += g`, one launch per parameter and pass; here it is the
 * last store of the producing kernel, so the gradients of a two-pass V2 step land in a flat gradient buffer with
 * no extra launches). */
enum { PP_FLAG_ACC_DW = 1, PP_FLAG_ACC_DGAMMA = 2, PP_FLAG_ACC_DBETA = 4,
       /* the HBM-bound passes of this backward call will share the SMs with a tensor-core kernel running on another
        * stream (the previous block's weight gradient): size their grids to two resident CTAs per SM, which still
        * saturates HBM and leaves registers / shared memory for that kernel's CTA, instead of every slot of an idle
        * chip (a statically partitioned grid that is only partly resident runs in two waves) */
       PP_FLAG_SHARE_SM = 8 };
enum { PP_WS_FWD = 0, PP_WS_BWD = 1 };
/* PPConvDesc.dtype — the arithmetic type of the contractions and the element type of every activation-shaped tensor
 * that crosses this boundary (x, y, dy, dx; the weight operand copies of pp_weight_prep):
 *   PP_DTYPE_BF16  bf16 tensors, tcgen05 kind::f16 (bf16 x bf16 -> fp32)               BASELINE configs 3-5
 *   PP_DTYPE_TF32  fp32 tensors read by the tensor cores as TF32, tcgen05 kind::tf32   BASELINE config 2
 *                  (what the reference's fp32 modules do on a GPU: torch leaves cudnn.allow_tf32 on; train_v1.py:13-29)
 * z (conv output saved for backward) is fp32 in TF32 mode; group / instance norm and the single-kernel passport block
 * are bf16-only (PP_EUNSUPPORTED / kernel sequence otherwise). */
enum { PP_DTYPE_BF16 = 0, PP_DTYPE_TF32 = 1 };

/* Geometry + mode of one conv block.  Mirrors the constructor arguments of the
 * reference blocks: PassportBlock(i, o, ks, s, pd, ...) models/layers/passportconv2d.py:12-18,
 * ConvBlock(i, o, ks, s, pd, bn, relu) models/layers/conv2d.py:6-9. */
typedef struct PPConvDesc {
  int32_t N, C, H, W; /* input activation, logical NCHW = [N,C,H,W], memory NHWC */
  int32_t O;          /* output channels                                         */
  int32_t kh, kw;     /* filter size                                             */
  int32_t stride;     /* same in h and w (reference passes one int)              */
  int32_t pad;        /* same in h and w                                         */
  int32_t norm;       /* PP_NORM_*  (bn in train / eval mode, group/instance norm, none) */
  int32_t relu;       /* 1: ReLU after the affine                                */
  int32_t z_f32;      /* 1: conv output z kept in fp32, 0: bf16                  */
  float eps;          /* BatchNorm eps (1e-5)                                    */
  float momentum;     /* BatchNorm momentum (0.1)                                */
  int32_t algo;       /* PP_ALGO_*; AUTO picks tcgen05 when C%64==0 && O%64==0   */
  int32_t groups;     /* PP_NORM_GN only: number of groups (O for InstanceNorm); else 0  */
  int32_t flags;      /* PP_FLAG_* (backward only); 0 = overwrite dw / dgamma / dbeta    */
  int32_t dtype;      /* PP_DTYPE_*: element type of x / y / dy / dx and of the weight operand copies */
} PPConvDesc;

int pp_version(void);
const char* pp_last_error(void);

/* Number of SMs / compute capability of the current device (for tests & bench). */
int pp_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* Bytes of scratch pp_conv_block_fwd (which=PP_WS_FWD) / _bwd (PP_WS_BWD) need. */
int pp_workspace_bytes(const PPConvDesc* d, int which, size_t* bytes);

/* fp32 OIHW master weight -> bf16 operand copies.
 *   w_fprop: [O, kh, kw, C]           (K-major B operand of the forward implicit GEMM)
 *   w_dgrad: [C, kh, kw, O]           (B operand of the data-gradient GEMM, taps not flipped; may be NULL)
 * Replaces the implicit weight read of nn.Conv2d (passportconv2d.py:18,218). */
int pp_weight_prep(const PPConvDesc* d, const float* w_oihw, void* w_fprop, void* w_dgrad, void* stream);

/* S[c*T + t] (T = kh*kw, t = r*kw + s: the element order of one row of the OIHW weight) = mean over key batch and
 * output positions of the (zero padded) key patch element (channel c, tap t); the fp32 key is NOT rounded,
 * accumulation is fp64.
 * With it  GAP(conv(W, key)) == W[O, kh*kw*C] @ S  (SURVEY 7.3), which replaces the two
 * batch-1 cuDNN convs + means of get_scale / get_bias (passportconv2d.py:146-152,167-173).
 * key_nchw: fp32 [Bk, C, H, W] exactly as the module's `key` / `skey` buffers store it. */
int pp_key_pool(const PPConvDesc* d, int Bk, const float* key_nchw, double* S, void* stream);

/* gamma = W @ S_skey, beta = W @ S_key with W the fp32 OIHW MASTER weight read as [O, C*kh*kw] (fp64 accumulate,
 * fp32 out) — not the bf16 operand copy: sign(gamma) is the embedded signature and has to be the one the
 * reference's fp32 get_scale() (passportconv2d.py:142-158) produces on the same weight and key.  Plus
 * SignLoss.add (models/losses/sign_loss.py:25-28,32-54):
 *   sign_loss = alpha * sum(relu(0.1 - b*gamma)) + 1e-5 * sum(gamma^2)
 *   sign_acc  = mean(sign(b) == sign(gamma))
 * b_sign may be NULL (then sign_loss/sign_acc are not written). */
int pp_passport_affine_fwd(const PPConvDesc* d, const float* w_oihw, const double* S_skey,
                           const double* S_key, const float* b_sign, float alpha, float* gamma,
                           float* beta, float* sign_loss, float* sign_acc, void* stream);

/* Gradient of the above w.r.t. the fp32 OIHW weight (rank-1 update):
 *   dW[o,c,r,s] = (g_gamma[o] + g_loss * dLsign/dgamma[o]) * S_skey[c,(r,s)] + g_beta[o] * S_key[c,(r,s)]
 * g_gamma / g_beta / g_loss may each be NULL (treated as 0); g_loss is a device scalar.
 * accumulate != 0 adds into dw_oihw instead of overwriting it. */
int pp_passport_affine_bwd(const PPConvDesc* d, const double* S_skey, const double* S_key,
                           const float* gamma, const float* b_sign, float alpha,
                           const float* g_gamma, const float* g_beta, const float* g_loss,
                           float* dw_oihw, int accumulate, void* stream);

/* Gradient of pp_passport_affine_fwd w.r.t. the passport keys themselves (needed when an attack turns `key` /
 * `skey` into Parameters: passport_attack_3.py:232-270).  d describes the geometry with H,W = key size.
 *   scratch: 2 * kh*kw*C doubles;  dskey_nchw / dkey_nchw: fp32 [Bk,C,H,W] (either may be NULL). */
int pp_passport_key_grad(const PPConvDesc* d, int Bk, const float* w_oihw, const float* gamma, const float* b_sign,
                         float alpha, const float* g_gamma, const float* g_beta, const float* g_loss,
                         double* scratch, float* dskey_nchw, float* dkey_nchw, void* stream);

/* Ownership verification of every passport layer of a model in ONE launch — the loop of
 * TesterPrivate.test_signature (experiments/trainer_private.py:37-71; also tester of passport_attack_1.py:112-170):
 *   signbit = m.get_scale(ind=1).view(-1).sign();  detection = (signbit == m.b).float().mean()
 * layers: HOST array of nlayers (<= PP_SIG_MAX_LAYERS) entries whose pointers are device pointers:
 *   w_oihw fp32 [O, C, kh, kw] (the master weight), S_skey fp64 [K] (pp_key_pool of skey), b_sign fp32 [O] (+-1),
 *   K = kh*kw*C, C = input channels.
 * matched: device int32 [nlayers], overwritten with the number of channels whose sign(gamma) equals b
 *   (detection = matched / O).  gamma_out: NULL or device fp32 buffer; layer i writes its gamma at
 *   gamma_out[layers[i].gamma_offset ...].  The arithmetic is that of pp_passport_affine_fwd, so the bits are
 *   identical to the per-layer path. */
#define PP_SIG_MAX_LAYERS 64
typedef struct PPSigLayer {
  const float* w_oihw;
  const double* S_skey;
  const float* b_sign;
  int32_t O;
  int32_t K;
  int32_t gamma_offset;
  int32_t C;
} PPSigLayer;
int pp_signature_verify(int nlayers, const PPSigLayer* layers, int32_t* matched, float* gamma_out, void* stream);

/* Stand-alone SignLoss.add on an arbitrary scale vector (sign_loss.py:18-54). */
int pp_sign_loss_fwd(int O, const float* gamma, const float* b_sign, float alpha, float* sign_loss,
                     float* sign_acc, void* stream);
int pp_sign_loss_bwd(int O, const float* gamma, const float* b_sign, float alpha, const float* g_loss,
                     float* g_gamma, void* stream);

/* y = relu?( gamma * norm(conv(x, W)) + beta )   — PassportBlock.forward
 * (passportconv2d.py:218-222), PassportPrivateBlock.forward (passportconv2d_private.py:215-218)
 * and ConvBlock.forward (conv2d.py:29-36, gamma/beta = BatchNorm affine).
 *   x        bf16 [N,H,W,C]          w_fprop   bf16 [O,kh,kw,C]
 *   gamma/beta fp32 [O]              running_* fp32 [O]  (updated in BN_TRAIN, read in BN_EVAL, may be NULL for NONE)
 *   z        conv output [N,P,Q,O] (bf16 or fp32 per d->z_f32); NULL => not kept (inference;
 *            then norm must not be BN_TRAIN / GN and the affine is fused into the conv epilogue)
 *   y        bf16 [N,P,Q,O]
 *   save_mean / save_invstd fp32 [O]: statistics the backward needs (batch stats, running
 *            stats, or 0/1 for NONE); fp32 [N*groups] (per sample and group) for PP_NORM_GN. */
int pp_conv_block_fwd(const PPConvDesc* d, const void* x, const void* w_fprop, const float* gamma,
                      const float* beta, float* running_mean, float* running_var, void* z, void* y,
                      float* save_mean, float* save_invstd, void* workspace, size_t ws_bytes,
                      void* stream);

/* The same block with the residual join of a ResNet basic unit folded into its last pass
 * (models/resnet_passport_private.py:78-85, resnet_passport.py:77-84, resnet_normal.py:24-26):
 *   y = bf16(relu(gamma * norm(conv(x, W)) + beta)) + residual        residual: bf16, same shape and layout as y
 * The reference computes relu(out + shortcut) where out and shortcut are both outputs of blocks that end in a ReLU
 * (every block of its ResNets does, convbn_2 and the shortcut included), so the outer ReLU is the identity; the caller
 * passes `residual` only when it is known to be >= 0 (a block / unit output).  The inner rounding makes the result
 * bit-identical to the separate add pass it replaces.  residual == NULL: exactly pp_conv_block_fwd.  bf16 tensors,
 * PP_NORM_NONE / PP_NORM_BN_* with the z buffer; other requests return PP_EUNSUPPORTED. */
int pp_conv_block_fwd_res(const PPConvDesc* d, const void* x, const void* w_fprop, const float* gamma,
                          const float* beta, float* running_mean, float* running_var, void* z, void* y,
                          float* save_mean, float* save_invstd, const void* residual, void* workspace,
                          size_t ws_bytes, void* stream);

/* Backward of pp_conv_block_fwd (autograd of the same reference lines).
 *   dy bf16 [N,P,Q,O];  dx bf16 [N,H,W,C] or NULL;  dw_oihw fp32 [O,C,kh,kw] or NULL
 *   dgamma / dbeta fp32 [O] (always written; added to when d->flags has PP_FLAG_ACC_DGAMMA / _DBETA,
 *   as dw_oihw is with PP_FLAG_ACC_DW). */
int pp_conv_block_bwd(const PPConvDesc* d, const void* dy, const void* x, const void* w_dgrad,
                      const void* z, const float* gamma, const float* beta, const float* save_mean,
                      const float* save_invstd, void* dx, float* dw_oihw, float* dgamma,
                      float* dbeta, void* workspace, size_t ws_bytes, void* stream);

/* THE PASSPORT BLOCK, forward — PassportBlock.forward (models/layers/passportconv2d.py:209-223) including get_scale /
 * get_bias (:142-175) and SignLoss.add (models/losses/sign_loss.py:32-54); PassportPrivateBlock.forward
 * (models/layers/passportconv2d_private.py:205-219) for both values of `ind`:
 *
 *   gamma = GAP(conv(W, skey)), beta = GAP(conv(W, key))     or the public scale / bias (ind == 0, no sign loss)
 *   sign_loss = alpha * sum(relu(0.1 - b * gamma)) + 1e-5 * sum(gamma^2);  sign_acc = mean(sign(b) == sign(gamma))
 *   y = relu?( gamma * bn(conv(x, W)) + beta ),  running statistics updated in BN_TRAIN
 *
 * Where every output tile of the convolution fits in tensor memory (O % 256 == 0 and ceil(N*P*Q / 128) * O / 256 <=
 * 2 * #SMs: ResNet-18 layer4 up to 1184 CIFAR / 386 ImageNet images per GPU), norm == PP_NORM_BN_TRAIN and z is kept
 * in fp32, this is ONE cooperative kernel: TMA-staged tiles, tcgen05 contractions with the accumulators resident in
 * TMEM, gamma / beta rows computed by the idle epilogue warps, per-CTA column statistics, a grid barrier, and
 * y = relu(a z + b) written straight from TMEM with 128-bit stores.  Otherwise the same result is produced by the
 * kernel sequence of pp_passport_affine_fwd + pp_conv_block_fwd.
 *
 *   w_fprop   bf16 operand copy (pp_weight_prep), w_oihw the fp32 master weight (gamma / beta are taken from it)
 *   S_skey / S_key  pooled keys (pp_key_pool); ignored when scale_pub / bias_pub are given (both or neither)
 *   gamma / beta    fp32 [O] outputs on the passport path (the backward needs gamma); untouched on the public path
 *   sign_loss / sign_acc  device scalars, written on the passport path when b_sign != NULL (may be NULL)
 *   other arguments as pp_conv_block_fwd. */
int pp_passport_conv_fwd(const PPConvDesc* d, const void* x, const void* w_fprop, const float* w_oihw,
                         const double* S_skey, const double* S_key, const float* scale_pub, const float* bias_pub,
                         const float* b_sign, float alpha, float* running_mean, float* running_var, void* y, void* z,
                         float* gamma, float* beta, float* save_mean, float* save_invstd, float* sign_loss,
                         float* sign_acc, void* workspace, size_t ws_bytes, void* stream);

/* THE PASSPORT BLOCK, backward (autograd of the lines above):
 *   dx = dgrad(W, dz);  dW = wgrad(x, dz) + (dgamma + g_sign_loss * dLsign/dgamma) (x) S_skey + dbeta (x) S_key
 *   with dLsign/dgamma_o = -alpha b_o [0.1 - b_o gamma_o > 0] + 2e-5 gamma_o,  dgamma = sum dy_m zhat, dbeta = sum dy_m.
 * Passport path: pass S_skey / S_key (and b_sign, alpha, the device scalar g_sign_loss = upstream gradient of the
 * sign loss, NULL = 0); dgamma / dbeta are scratch outputs.  Public path (S_skey == S_key == NULL): gamma / beta are
 * the public scale / bias and dgamma / dbeta their gradients (dscale, dbias).  PP_FLAG_ACC_DW in d->flags adds into
 * dw_oihw.  Other arguments as pp_conv_block_bwd. */
int pp_passport_conv_bwd(const PPConvDesc* d, const void* dy, const void* x, const void* w_dgrad, const void* z,
                         const float* gamma, const float* beta, const float* save_mean, const float* save_invstd,
                         const double* S_skey, const double* S_key, const float* b_sign, float alpha,
                         const float* g_sign_loss, void* dx, float* dw_oihw, float* dgamma, float* dbeta,
                         void* workspace, size_t ws_bytes, void* stream);

/* The block backward WITHOUT its weight gradient: dgamma, dbeta, dx as pp_conv_block_bwd, and dz (the gradient w.r.t.
 * the conv output, [N*P*Q, O] in the activation type) written to the caller's buffer dz_out instead of the workspace.
 * The weight gradient is then pp_conv_wgrad(d, dz_out, x, dw, ...) — a tensor-core kernel nothing downstream in the
 * backward pass depends on, which the caller may launch on a second stream so that it overlaps the HBM-bound passes
 * of the next block's backward (deepipr_b200.functional does, into a flat gradient buffer with PP_FLAG_ACC_DW).
 * Batch-norm / plain blocks only (PP_NORM_GN: PP_EUNSUPPORTED).
 * dx_add (optional, bf16, the shape of dx): a gradient that reaches the block's input by another route (the residual
 * path of a ResNet unit, resnet_passport_private.py:78-85) and is added to the data gradient in the dgrad kernel's
 * epilogue, dx = bf16(bf16(dgrad) + dx_add) — the sum autograd would otherwise form in a separate pass. */
int pp_conv_block_bwd_dz(const PPConvDesc* d, const void* dy, const void* w_dgrad, const void* z, const float* gamma,
                         const float* beta, const float* save_mean, const float* save_invstd, void* dx,
                         const void* dx_add, float* dgamma, float* dbeta, void* dz_out, void* workspace,
                         size_t ws_bytes, void* stream);

/* Building blocks exposed for tests / profiling (same kernels the two calls above use).  pp_conv_wgrad honours
 * PP_FLAG_ACC_DW in d->flags (adds into dw_oihw). */
int pp_conv_fwd_raw(const PPConvDesc* d, const void* x, const void* w_fprop, void* z, void* workspace,
                    size_t ws_bytes, void* stream); /* z = conv(x,W), dtype per z_f32 */
int pp_conv_dgrad(const PPConvDesc* d, const void* dz, const void* w_dgrad, void* dx, void* stream);
int pp_conv_wgrad(const PPConvDesc* d, const void* dz, const void* x, float* dw_oihw, void* workspace,
                  size_t ws_bytes, void* stream);

/* Residual join of a ResNet basic block (models/resnet_passport_private.py:78-85, resnet_passport.py:77-84,
 * resnet_normal.py:24-26): y = relu(a + b) on bf16 tensors of n elements, and its backward
 * gx = gy * [y > 0] (the same gradient flows to both inputs). */
int pp_add_relu_fwd(size_t n, const void* a, const void* b, void* y, void* stream);
int pp_add_relu_bwd(size_t n, const void* gy, const void* y, void* gx, void* stream);

/* nn.MaxPool2d(k, stride, pad) (dilation 1, floor mode) on a dense NHWC tensor, bf16 or fp32 (f32 != 0): the pooling
 * layers between the blocks of the reference nets (models/resnet_passport.py / resnet_passport_private.py:
 * MaxPool2d(3, 2, 1) behind the ImageNet stem; models/alexnet_passport.py:37-38: MaxPool2d(2, 2)).
 *   y[n,ph,pw,c] = max over the window; argmax[n,ph,pw,c] = position r*k + t of the FIRST maximum in scan order
 *   (ATen's rule), one byte per output element;  dx = gather of dy through argmax (no atomics, deterministic).
 * C % 8 == 0.  x: [N,H,W,C], y / argmax / dy: [N,P,Q,C] with P = (H + 2 pad - k) / stride + 1, dx: [N,H,W,C]. */
int pp_maxpool_fwd(int N, int H, int W, int C, int k, int stride, int pad, const void* x, int f32, void* y,
                   uint8_t* argmax, void* stream);
int pp_maxpool_bwd(int N, int H, int W, int C, int k, int stride, int pad, const void* dy, const uint8_t* argmax,
                   int f32, void* dx, void* stream);

/* Fused SGD(momentum, weight decay) step on one flat fp32 buffer (classification.py:47-50):
 *   g' = g + wd*p;  buf = mom*buf + g' (buf = g' on the first step);  p -= lr*buf */
int pp_sgd_step(size_t n, float* param, const float* grad, float* momentum_buf, float lr, float momentum,
                float weight_decay, int first_step, void* stream);

/* The same update with {lr, momentum, weight_decay, first_step} read from DEVICE memory (`hyper`, 4 floats): a
 * CUDA-graph-captured step keeps following a learning-rate schedule without being re-captured. */
int pp_sgd_step_dev(size_t n, float* param, const float* grad, float* momentum_buf, const float* hyper,
                    void* stream);

/* Loss / metric epilogue of one forward pass in ONE launch, forward and backward
 * (experiments/trainer_private.py:161-168, experiments/trainer.py:28-43,139-142):
 *   loss[0]  (+)= mean_n( logsumexp(logits[n,:]) - logits[n,target[n]] )         F.cross_entropy(pred, target)
 *   top1[0]  (+)= 100/N * #{n : argmax_c logits[n,c] == target[n]}               accuracy(pred, target)[0]
 *   dlogits[n,c] = (softmax(logits[n,:])[c] - [c == target[n]]) / N              d(mean loss)/d(logits), fp32
 * logits: [N, classes] fp32 or bf16 (logits_bf16 != 0), row-major; target: int64 [N]; loss / top1 / dlogits may each
 * be NULL; accumulate != 0 adds into loss / top1 (the V2 loop sums the loss of its two passes).  With it the trainer
 * step needs no host read at all: the four numbers the reference prints are device scalars read once per step or
 * per epoch. */
int pp_ce_top1(int N, int classes, const void* logits, int logits_bf16, const int64_t* target, float* loss,
               float* top1, float* dlogits, int accumulate, void* stream);

/* Debug: after a kernel-side pipeline timeout the offending barrier id is recorded here. */
int pp_debug_last_timeout(void);
/* A/B switch for tests and profiling: on = 0 makes every block take the kernel SEQUENCE even where the single
 * cooperative kernel applies, on = 1 re-enables it, on < 0 only queries.  Returns the previous setting. */
int pp_debug_fused(int on);

/* Instrumentation used by bench.py.
 *   pp_launch_count   kernels launched by this library since the last reset (the "gpu_launches" claim)
 *   pp_profile_*      when enabled, every tensor-core kernel launch is bracketed by CUDA events on its stream;
 *                     pp_profile_read(kind, ...) sums durations and algorithmic FLOPs of the recorded launches
 *                     (kind 0: implicit-GEMM conv / dgrad kernel, kind 1: weight-gradient kernel), optionally
 *                     only those whose GEMM has `filter_c` input channels, `filter_nout` output columns and
 *                     `filter_taps` taps (<= 0: any).  Enabling starts a fresh recording. */
long long pp_launch_count(int reset);
int pp_profile_enable(int on);
int pp_profile_read(int kind, int filter_c, int filter_nout, int filter_taps, double* total_ms,
                    double* total_flops, int* launches);

#ifdef __cplusplus
}
#endif
#endif /* PASSPORT_SM100_H_ */
