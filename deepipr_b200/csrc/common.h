// common.h — host-side shared declarations of libpassport_sm100 (error reporting, launch geometry).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/passport_sm100.h"

namespace pp {

// thread-local last error message (pp_last_error)
void set_error(const char* fmt, ...);
const char* get_error();

#define PP_CHECK_CUDA(expr)                                                                     \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      pp::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));      \
      return PP_ELAUNCH;                                                                        \
    }                                                                                           \
  } while (0)

// after every kernel launch: count it (pp_launch_count) and surface launch errors
#define PP_POST_LAUNCH()                  \
  do {                                    \
    pp::count_launch();                   \
    PP_CHECK_CUDA(cudaGetLastError());    \
  } while (0)

#define PP_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      pp::set_error(__VA_ARGS__);    \
      return (code);                 \
    }                                \
  } while (0)

// cudaFuncSetAttribute is a per-DEVICE setting: remember it per device ordinal, not per process (a process may
// drive several GPUs, e.g. under nn.DataParallel).  `kernel` is passed in parentheses (template commas).
constexpr int kMaxDevices = 64;
#define PP_SET_MAX_SMEM_ONCE(kernel, bytes)                                                                  \
  do {                                                                                                      \
    static bool done_[pp::kMaxDevices] = {};                                                                \
    const int dev_ = pp::current_device();                                                                  \
    if (dev_ < 0 || dev_ >= pp::kMaxDevices || !done_[dev_]) {                                              \
      PP_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (bytes)));   \
      if (dev_ >= 0 && dev_ < pp::kMaxDevices) done_[dev_] = true;                                          \
    }                                                                                                       \
  } while (0)

#define PP_TRY(expr)            \
  do {                          \
    int _s = (expr);            \
    if (_s != PP_OK) return _s; \
  } while (0)

constexpr int kMaxTaps = 49;  // up to 7x7 filters

// "Tap-list implicit GEMM":  D[m, n] = sum_t sum_c act[pixel(m) + (dh_t, dw_t), c] * B[n, kofs_t + c]
//   m enumerates a traversal grid of P x Q positions per image (q fastest, then p, then image):
//     act pixel = (base_h + p*step_h + dh, base_w + q*step_w + dw), zero outside the tensor
//   output row m is written at pixel (p*out_sh + out_ph, q*out_sw + out_pw) of an out_H x out_W image.
// The forward conv, the stride-1 data gradient and every phase of a strided data gradient are all
// instances of this (see api.cu: plan_fprop / plan_dgrad_phase / plan_col).
struct TapGemm {
  // activation tensor, NHWC bf16 — or NHWC fp32 read as TF32 when tf32 != 0 (PP_DTYPE_TF32; B is fp32 too then)
  int tf32;
  int N, H, W, C;
  // traversal grid
  int P, Q;
  int base_h, base_w, step_h, step_w;
  int upper_h, upper_w;  // im2col bounding-box upper corner (derived from P,Q; kept for the TMA descriptor)
  int ntaps;
  int8_t tap_dh[kMaxTaps], tap_dw[kMaxTaps];
  int tap_kofs[kMaxTaps];
  // B operand: [Nout, Ktot] row-major bf16
  int Nout, Ktot;
  // output mapping
  int out_H, out_W, out_sh, out_sw, out_ph, out_pw;
  int out_identity;  // 1: out pixel == m (plain [M, Nout] matrix)
};

// Epilogue of the tap GEMM
struct TapEpilogue {
  void* out;             // [out pixels, Nout], bf16 or fp32
  int out_f32;           // 1: fp32
  const float* scale;    // per-n affine a[n] (NULL => 1)
  const float* shift;    // per-n affine b[n] (NULL => 0)
  int relu;
  float* stats_partial;  // NULL or [tapgemm_tcgen05_grid()][2][Nout]: per-CTA sum z, sum z^2 (fp32 accumulators)
  // NULL, or a bf16 tensor of the output's shape and pixel mapping that is added to the (bf16-rounded) result before
  // it is stored: out = bf16(bf16(acc) + add_src) — the gradient sum autograd performs where a tensor forks
  // (dx = dgrad + gradient of the residual path), folded into the data-gradient kernel.  bf16 outputs only.
  const void* add_src = nullptr;
};

void count_launch();
// optional per-kernel CUDA-event timing (bench.py roofline legs): the two tensor-core kernels (work = FLOPs) and the
// three HBM-bound passes of the block (work = algorithmic bytes: affine z->y, backward reduce, dz)
enum { PROF_TAPGEMM = 0, PROF_WGRAD = 1, PROF_AFFINE = 2, PROF_REDUCE = 3, PROF_DZ = 4, PROF_FUSED = 5, PROF_KINDS = 6 };
void prof_begin(int kind, double flops, int c, int nout, int taps, cudaStream_t s);
void prof_end(int kind, cudaStream_t s);

int current_device();   // ordinal of the calling thread's current CUDA device (-1: none)
int device_sm_count();
int check_device();  // PP_OK iff current device is sm_100

// --- igemm_sm100.cu (tcgen05) ---
int tapgemm_tcgen05(const TapGemm& g, const void* act, const void* B, const TapEpilogue& e, cudaStream_t s);
bool tapgemm_tcgen05_supported(const TapGemm& g);
int tapgemm_tcgen05_grid(const TapGemm& g);  // CTAs launched == rows of stats_partial written
int tapgemm_tcgen05_max_stats_width();       // widest Nout for which the fused column statistics are available
int debug_last_timeout();
// the passport block as one cooperative kernel (conv + batch statistics + grid barrier + affine + ReLU from TMEM)
struct FusedArgs {
  void* y;                       // bf16 [rows, O]
  float* z;                      // fp32 [rows, O] (saved for backward)
  float* partial;                // [passport_fused_grid()][2][256]
  unsigned int* barrier;         // 4 bytes of workspace
  const float* gamma_in; const float* beta_in;     // given per-channel affine, or ...
  const float* w_oihw; const double* Ss; const double* Sk; int Cin, T;   // ... derived from the passport
  float* gamma_out; float* beta_out;
  const float* b_sign; float alpha; float* sign_loss; float* sign_acc;
  float* rmean; float* rvar; float* save_mean; float* save_invstd;
  float eps, momentum;
  int relu;
};
int debug_fused(int on);   // A/B switch of the single-kernel block: returns the previous setting, on < 0 only queries
bool passport_fused_supported(const TapGemm& g);
int passport_fused_grid(const TapGemm& g);
int passport_fused_tcgen05(const TapGemm& g, const void* act, const void* B, const FusedArgs& a, cudaStream_t s);
// wgrad: partial[split][Mo][T*C] (fp32) = sum over a slice of pixels of dz[m, o] * act_tap[m, c]
int wgrad_tcgen05(const TapGemm& g /*fprop geometry of x*/, const void* x, const void* dz, int O, float* partial,
                  int splits, cudaStream_t s);
bool wgrad_tcgen05_supported(const TapGemm& g, int O);
int wgrad_pick_splits(const TapGemm& g, int O);

// --- direct_conv.cu (SIMT, any shape) ---
int tapgemm_simt(const TapGemm& g, const void* act, const void* B, const TapEpilogue& e, cudaStream_t s);
int wgrad_simt(const TapGemm& g, const void* x, const void* dz, int O, float* partial, int splits, cudaStream_t s);
int wgrad_simt_pick_splits(const TapGemm& g, int O);

// --- stem_conv.cu (direct CUDA-core kernels for the 3 -> 64 channel 3x3/s1/p1 stem) ---
bool stem_direct_supported(const PPConvDesc& d);
int stem_grid(const PPConvDesc& d);  // CTAs launched == rows of stats_partial / wgrad partial tiles written
int stem_fprop(const PPConvDesc& d, const void* x, const void* wf, const TapEpilogue& e, cudaStream_t s);
int stem_wgrad(const PPConvDesc& d, const void* x, const void* dz, float* partial /*[grid][64][27]*/, cudaStream_t s);

// --- pointwise.cu ---
int launch_weight_prep(const PPConvDesc& d, const float* w, void* wf, void* wd, cudaStream_t s);  // element type per d.dtype
int launch_key_pool(const PPConvDesc& d, int Bk, const float* key, double* S, cudaStream_t s);
int launch_passport_affine_fwd(const PPConvDesc& d, const float* w_oihw, const double* Ss, const double* Sk,
                               const float* b, float alpha, float* gamma, float* beta, float* loss, float* acc,
                               cudaStream_t s);
int launch_passport_affine_bwd(const PPConvDesc& d, const double* Ss, const double* Sk, const float* gamma,
                               const float* b, float alpha, const float* gg, const float* gb, const float* gl,
                               float* dw, int accumulate, cudaStream_t s);
int launch_passport_key_grad(const PPConvDesc& d, int Bk, const float* w_oihw, const float* gamma, const float* b,
                             float alpha, const float* gg, const float* gb, const float* gl, double* dSs, double* dSk,
                             float* dskey, float* dkey, cudaStream_t s);
int launch_signature_verify(int nlayers, const PPSigLayer* layers, int* matched, float* gamma_out, cudaStream_t s);
int launch_sign_loss_fwd(int O, const float* gamma, const float* b, float alpha, float* loss, float* acc,
                         cudaStream_t s);
int launch_sign_loss_bwd(int O, const float* gamma, const float* b, float alpha, const float* gl, float* gg,
                         cudaStream_t s);
// (a, b) affine coefficient vectors + statistics
int launch_bn_finalize(const PPConvDesc& d, int n_per_channel, const float* stats_partial, int num_tiles,
                       const float* gamma, const float* beta, float* running_mean, float* running_var,
                       float* save_mean, float* save_invstd, float* coef_a, float* coef_b, cudaStream_t s);
int launch_affine_apply(const void* z, int z_f32, size_t rows, int O, const float* a, const float* b, int relu,
                        void* y, int y_f32, const void* res /*bf16 residual added after the ReLU, or NULL*/,
                        cudaStream_t s);
int launch_col_stats(const void* z, int z_f32, size_t rows, int O, float* partial, int* num_partials,
                     cudaStream_t s);
int launch_bwd_reduce(const void* dy, int dy_f32, const void* z, int z_f32, size_t rows, int O, const float* gamma,
                      const float* beta, const float* mean, const float* invstd, int relu, float* partial,
                      int* num_partials, cudaStream_t s, int share_sm = 0);   // mask coefficients derived in the kernel
int bwd_reduce_max_partials();
int launch_bwd_coef(const PPConvDesc& d, size_t rows, const float* partial, int num_partials, const float* gamma,
                    const float* beta, const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta,
                    float* k1, float* k2, float* k3, float* ca, float* cb /*mask coefficients for the dz pass*/,
                    cudaStream_t s);
int launch_bwd_dz(const void* dy, int act_f32 /*dy and dz*/, const void* z, int z_f32, size_t rows, int O, const float* a,
                  const float* b, int relu, const float* k1, const float* k2, const float* k3, void* dz,
                  cudaStream_t s, int share_sm = 0);
int launch_wgrad_finalize(const PPConvDesc& d, const float* partial, int splits, int kstride, float* dw_oihw,
                          cudaStream_t s, int accumulate = 0);
int launch_im2col_small(const PPConvDesc& d, const void* x, void* col, size_t rows, int P, int Q, int Kpad,
                        cudaStream_t s);   // element type per d.dtype
int launch_pad_rows(const void* src, void* dst, int rows, int K, int Kpad, int f32, cudaStream_t s);
int launch_maxpool_fwd(int N, int H, int W, int C, int k, int s, int p, const void* x, int f32, void* y, uint8_t* idx,
                       cudaStream_t st);
int launch_maxpool_bwd(int N, int H, int W, int C, int k, int s, int p, const void* dy, const uint8_t* idx, int f32,
                       void* dx, cudaStream_t st);
int launch_add_inplace(__nv_bfloat16* y, const __nv_bfloat16* a, size_t n, cudaStream_t s);   // y += a
int launch_add_relu_fwd(const __nv_bfloat16* a, const __nv_bfloat16* b, __nv_bfloat16* y, size_t n, cudaStream_t s);
int launch_add_relu_bwd(const __nv_bfloat16* g, const __nv_bfloat16* y, __nv_bfloat16* gx, size_t n, cudaStream_t s);
// --- groupnorm.cu (GroupNorm / InstanceNorm: per-(sample, group) statistics, per-(sample, channel) coefficients) ---
int gn_chunks(int N, int HW, int O);  // partial rows per sample written by the statistics kernels
int launch_gn_fwd(const PPConvDesc& d, int HW, const void* z, const float* gamma, const float* beta, float* save_mean,
                  float* save_invstd, float* ca, float* cb, float* partial, __nv_bfloat16* y, cudaStream_t s);
int launch_gn_bwd_reduce(const PPConvDesc& d, int HW, const __nv_bfloat16* dy, const void* z, const float* gamma,
                         const float* beta, const float* save_mean, const float* save_invstd, float* ca, float* cb,
                         float* k1, float* k2, float* k3, float* partial, float* contrib, float* dgamma, float* dbeta,
                         cudaStream_t s);
int launch_gn_dz(const PPConvDesc& d, int HW, const __nv_bfloat16* dy, const void* z, const float* ca,
                 const float* cb, const float* k1, const float* k2, const float* k3, __nv_bfloat16* dz,
                 cudaStream_t s);
int launch_sgd(size_t n, float* p, const float* g, float* buf, float lr, float mom, float wd, int first,
               cudaStream_t s);
int launch_sgd_dev(size_t n, float* p, const float* g, float* buf, const float* hyper, cudaStream_t s);
int launch_ce_top1(int N, int classes, const void* logits, int logits_bf16, const long long* target, float* loss,
                   float* top1, float* dlogits, int accumulate, cudaStream_t s);

}  // namespace pp
