// api.cu — extern "C" entry points of libpassport_sm100.so (declared in include/passport_sm100.h):
// argument checking, geometry planning (conv -> tap-GEMM instances) and kernel sequencing.
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.h"

namespace pp {

// ---------------------------------------------------------------- error + device
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

// per-device attributes (a process may drive several devices; cached by ordinal)
struct DevInfo { int sm = -1, cc_major = -1, cc_minor = -1; };
static DevInfo g_dev[kMaxDevices];
static std::mutex g_dev_mu;

int current_device() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return dev;
}

// ---------------------------------------------------------------- launch counter + kernel profiling
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct ProfRec { cudaEvent_t start, stop; double flops; int c, nout, taps; };
static bool g_prof_on = false;
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof[PROF_KINDS];

void prof_begin(int kind, double flops, int c, int nout, int taps, cudaStream_t s) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r;
  r.flops = flops; r.c = c; r.nout = nout; r.taps = taps;
  cudaEventCreate(&r.start);
  cudaEventCreate(&r.stop);
  cudaEventRecord(r.start, s);
  g_prof[kind].push_back(r);
}
void prof_end(int kind, cudaStream_t s) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof[kind].empty()) cudaEventRecord(g_prof[kind].back().stop, s);
}

static int query_device(DevInfo* out) {
  const int dev = current_device();
  if (dev < 0) {
    set_error("no CUDA device (libpassport_sm100 has no CPU path)");
    return PP_ENODEVICE;
  }
  if (dev < kMaxDevices) {
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (g_dev[dev].sm >= 0) { *out = g_dev[dev]; return PP_OK; }
  }
  DevInfo di;
  if (cudaDeviceGetAttribute(&di.sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&di.cc_major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&di.cc_minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) {
    cudaGetLastError();
    set_error("cannot query CUDA device attributes");
    return PP_ENODEVICE;
  }
  if (dev < kMaxDevices) {
    std::lock_guard<std::mutex> lk(g_dev_mu);
    g_dev[dev] = di;
  }
  *out = di;
  return PP_OK;
}

int device_sm_count() {
  DevInfo di;
  if (query_device(&di) != PP_OK) return 0;
  return di.sm;
}

int check_device() {
  DevInfo di;
  PP_TRY(query_device(&di));
  PP_REQUIRE(di.cc_major == 10, PP_ENODEVICE, "device is sm_%d%d; this library is built for sm_100a only", di.cc_major,
             di.cc_minor);
  return PP_OK;
}

// ---------------------------------------------------------------- geometry
struct Geo {
  int P, Q, T;
  size_t rows;  // N*P*Q
};

static int geo_of(const PPConvDesc* d, Geo* g) {
  PP_REQUIRE(d != nullptr, PP_EBADARG, "PPConvDesc is NULL");
  PP_REQUIRE(d->N > 0 && d->C > 0 && d->H > 0 && d->W > 0 && d->O > 0 && d->kh > 0 && d->kw > 0 && d->stride > 0 &&
                 d->pad >= 0,
             PP_EBADSHAPE, "bad conv desc N=%d C=%d H=%d W=%d O=%d k=%dx%d s=%d p=%d", d->N, d->C, d->H, d->W, d->O,
             d->kh, d->kw, d->stride, d->pad);
  PP_REQUIRE(d->kh * d->kw <= kMaxTaps, PP_EUNSUPPORTED, "filter %dx%d larger than %d taps", d->kh, d->kw, kMaxTaps);
  PP_REQUIRE(d->H + 2 * d->pad >= d->kh && d->W + 2 * d->pad >= d->kw, PP_EBADSHAPE, "filter larger than padded input");
  g->P = (d->H + 2 * d->pad - d->kh) / d->stride + 1;
  g->Q = (d->W + 2 * d->pad - d->kw) / d->stride + 1;
  g->T = d->kh * d->kw;
  g->rows = (size_t)d->N * g->P * g->Q;
  PP_REQUIRE(g->rows < (1ull << 31), PP_EBADSHAPE, "too many output pixels");
  PP_REQUIRE(d->dtype == PP_DTYPE_BF16 || d->dtype == PP_DTYPE_TF32, PP_EBADARG, "unknown dtype %d", d->dtype);
  PP_REQUIRE(d->dtype != PP_DTYPE_TF32 || d->norm != PP_NORM_GN, PP_EUNSUPPORTED,
             "group / instance norm blocks are bf16-only (PP_DTYPE_TF32 requested)");
  PP_REQUIRE(d->dtype != PP_DTYPE_TF32 || d->algo != PP_ALGO_SIMT, PP_EUNSUPPORTED,
             "the SIMT reference kernels are bf16-only (PP_DTYPE_TF32 requested)");
  return PP_OK;
}

static inline bool is_tf32(const PPConvDesc& d) { return d.dtype == PP_DTYPE_TF32; }
static inline size_t act_esz(const PPConvDesc& d) { return is_tf32(d) ? 4 : 2; }   // bytes per x / y / dy / dx / dz element

// forward conv as a tap-GEMM over x (models/layers/passportconv2d.py:18,218)
static void plan_fprop(const PPConvDesc& d, const Geo& geo, TapGemm& g) {
  memset(&g, 0, sizeof(g));
  g.tf32 = is_tf32(d) ? 1 : 0;
  g.N = d.N; g.H = d.H; g.W = d.W; g.C = d.C;
  g.P = geo.P; g.Q = geo.Q;
  g.base_h = -d.pad; g.base_w = -d.pad;
  g.step_h = d.stride; g.step_w = d.stride;
  g.upper_h = d.pad - (d.kh - 1);
  g.upper_w = d.pad - (d.kw - 1);
  g.ntaps = geo.T;
  for (int r = 0; r < d.kh; ++r)
    for (int s = 0; s < d.kw; ++s) {
      const int t = r * d.kw + s;
      g.tap_dh[t] = (int8_t)r; g.tap_dw[t] = (int8_t)s; g.tap_kofs[t] = t * d.C;
    }
  g.Nout = d.O; g.Ktot = geo.T * d.C;
  g.out_H = geo.P; g.out_W = geo.Q; g.out_sh = 1; g.out_sw = 1; g.out_ph = 0; g.out_pw = 0; g.out_identity = 1;
}

// data gradient: dx[h] = sum_r dz[(h + pad - r)/s] W[r] over the r with (h + pad - r) % s == 0.
// One tap-GEMM over dz per output phase (h % s, w % s); returns 0 if the phase has no taps (dx there is 0).
static int plan_dgrad_phase(const PPConvDesc& d, const Geo& geo, int ph, int pw, TapGemm& g) {
  memset(&g, 0, sizeof(g));
  const int s = d.stride;
  const int Hph = (d.H - ph + s - 1) / s;
  const int Wpw = (d.W - pw + s - 1) / s;
  if (Hph <= 0 || Wpw <= 0) return 0;
  int rs[8], es_h[8], nh = 0, ss[8], es_w[8], nw = 0;
  for (int r = 0; r < d.kh; ++r) {
    const int v = ph + d.pad - r;
    if (((v % s) + s) % s == 0) { rs[nh] = r; es_h[nh] = (v >= 0) ? v / s : -((-v) / s); ++nh; }
  }
  for (int c = 0; c < d.kw; ++c) {
    const int v = pw + d.pad - c;
    if (((v % s) + s) % s == 0) { ss[nw] = c; es_w[nw] = (v >= 0) ? v / s : -((-v) / s); ++nw; }
  }
  if (nh == 0 || nw == 0) return 0;
  int bh = es_h[0], bw = es_w[0];
  for (int i = 1; i < nh; ++i) bh = es_h[i] < bh ? es_h[i] : bh;
  for (int i = 1; i < nw; ++i) bw = es_w[i] < bw ? es_w[i] : bw;
  g.tf32 = is_tf32(d) ? 1 : 0;
  g.N = d.N; g.H = geo.P; g.W = geo.Q; g.C = d.O;  // activation of this GEMM is dz
  g.P = Hph; g.Q = Wpw;
  g.base_h = bh; g.base_w = bw; g.step_h = 1; g.step_w = 1;
  g.upper_h = Hph - geo.P + bh;
  g.upper_w = Wpw - geo.Q + bw;
  g.ntaps = 0;
  for (int i = 0; i < nh; ++i)
    for (int j = 0; j < nw; ++j) {
      const int t = g.ntaps++;
      g.tap_dh[t] = (int8_t)(es_h[i] - bh);
      g.tap_dw[t] = (int8_t)(es_w[j] - bw);
      g.tap_kofs[t] = (rs[i] * d.kw + ss[j]) * d.O;  // w_dgrad is [C][kh*kw][O]
    }
  g.Nout = d.C; g.Ktot = geo.T * d.O;
  g.out_H = d.H; g.out_W = d.W; g.out_sh = s; g.out_sw = s; g.out_ph = ph; g.out_pw = pw;
  g.out_identity = (s == 1) ? 1 : 0;
  return 1;
}

static bool use_tcgen05(const PPConvDesc& d, bool supported) {
  if (d.algo == PP_ALGO_SIMT) return false;
  return supported;
}

static inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// Small-C convolutions (C*kh*kw padded to Kpad, e.g. the 3-channel stem): explicit im2col + 1x1 tap-GEMM.
static int col_kpad(const PPConvDesc& d, const Geo& geo) {
  if (d.algo == PP_ALGO_SIMT) return 0;
  const int kq = is_tf32(d) ? 32 : 64;   // elements per 128-byte operand row
  if (d.C % kq == 0 || d.O % 64 != 0) return 0;
  const int K = d.C * geo.T;
  if (K > 1024) return 0;
  return (K + kq - 1) / kq * kq;
}

// the explicitly im2col'ed matrix col[rows][Kpad] viewed as a 1 x rows image with Kpad channels
static void plan_col(const PPConvDesc& d, const Geo& geo, int Kpad, TapGemm& g) {
  memset(&g, 0, sizeof(g));
  g.tf32 = is_tf32(d) ? 1 : 0;
  g.N = 1; g.H = 1; g.W = (int)geo.rows; g.C = Kpad;
  g.P = 1; g.Q = (int)geo.rows;
  g.base_h = 0; g.base_w = 0; g.step_h = 1; g.step_w = 1; g.upper_h = 0; g.upper_w = 0;
  g.ntaps = 1; g.tap_dh[0] = 0; g.tap_dw[0] = 0; g.tap_kofs[0] = 0;
  g.Nout = d.O; g.Ktot = Kpad;
  g.out_H = 1; g.out_W = (int)geo.rows; g.out_sh = 1; g.out_sw = 1; g.out_ph = 0; g.out_pw = 0; g.out_identity = 1;
}

// rows of partial[.][2][O] the GroupNorm statistics kernels write (groupnorm.cu)
static size_t gn_partial_rows(const PPConvDesc& d, const Geo& geo) {
  if (d.O % 8 != 0 || d.O > 2048) return 1;  // rejected later by the launch; keep the carve well-defined
  return (size_t)d.N * gn_chunks(d.N, geo.P * geo.Q, d.O);
}

struct FwdWs {
  float* ca; float* cb; float* partial;
  void* col; void* wpad;   // small-C im2col scratch (element type per d.dtype)
  unsigned int* barrier;   // grid-barrier counter of the single-kernel block
  size_t total;
};
static FwdWs carve_fwd(const PPConvDesc& d, const Geo& geo, void* base) {
  FwdWs w;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  size_t off = 0;
  const size_t coef_rows = d.norm == PP_NORM_GN ? (size_t)d.N : 1;  // GN/IN: a coefficient row per sample
  w.ca = reinterpret_cast<float*>(p + off); off += align256(coef_rows * d.O * 4);
  w.cb = reinterpret_cast<float*>(p + off); off += align256(coef_rows * d.O * 4);
  size_t max_part = (size_t)(bwd_reduce_max_partials() > 160 ? bwd_reduce_max_partials() : 160);
  if (d.norm == PP_NORM_GN) max_part = gn_partial_rows(d, geo);
  w.partial = reinterpret_cast<float*>(p + off); off += align256(max_part * 2 * d.O * 4);
  const int Kpad = col_kpad(d, geo);
  w.col = p + off; off += align256(geo.rows * (size_t)Kpad * act_esz(d));
  w.wpad = p + off; off += align256((size_t)d.O * Kpad * act_esz(d));
  w.barrier = reinterpret_cast<unsigned int*>(p + off); off += 256;
  w.total = off;
  return w;
}

struct BwdWs {
  float* ca; float* cb; float* k1; float* k2; float* k3; float* partial; float* contrib;
  void* dz; void* col; float* wpartial;   // dz / col: element type per d.dtype
  size_t total;
};
static BwdWs carve_bwd(const PPConvDesc& d, const Geo& geo, void* base, int wg_splits) {
  BwdWs w;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  size_t off = 0;
  const bool gn = d.norm == PP_NORM_GN;
  const size_t vec = align256((gn ? (size_t)d.N : 1) * d.O * 4);
  w.ca = reinterpret_cast<float*>(p + off); off += vec;
  w.cb = reinterpret_cast<float*>(p + off); off += vec;
  w.k1 = reinterpret_cast<float*>(p + off); off += vec;
  w.k2 = reinterpret_cast<float*>(p + off); off += vec;
  w.k3 = reinterpret_cast<float*>(p + off); off += vec;
  const size_t part_rows = gn ? gn_partial_rows(d, geo) : (size_t)bwd_reduce_max_partials();
  w.partial = reinterpret_cast<float*>(p + off); off += align256(part_rows * 2 * d.O * 4);
  w.contrib = reinterpret_cast<float*>(p + off); off += gn ? align256((size_t)d.N * 2 * d.O * 4) : 0;
  w.dz = p + off; off += align256(geo.rows * d.O * act_esz(d));
  const int Kpad = col_kpad(d, geo);
  w.col = p + off; off += align256(geo.rows * (size_t)Kpad * act_esz(d));
  const size_t krow = Kpad ? (size_t)Kpad : (size_t)geo.T * d.C;
  w.wpartial = reinterpret_cast<float*>(p + off);
  off += align256((size_t)wg_splits * d.O * krow * 4);
  w.total = off;
  return w;
}

static int wgrad_splits_for(const PPConvDesc& d, const Geo& geo, bool* tc) {
  TapGemm g;
  if (stem_direct_supported(d)) {   // one partial tile per CTA of stem_wgrad_kernel
    *tc = false;
    return stem_grid(d);
  }
  const int Kpad = col_kpad(d, geo);
  if (Kpad) {
    plan_col(d, geo, Kpad, g);
    *tc = true;
    return wgrad_pick_splits(g, d.O);
  }
  plan_fprop(d, geo, g);
  *tc = use_tcgen05(d, wgrad_tcgen05_supported(g, d.O));
  return *tc ? wgrad_pick_splits(g, d.O) : wgrad_simt_pick_splits(g, d.O);
}

// scratch: col / wpad are only touched on the small-C im2col path (may be NULL otherwise)
// *stats_rows (optional) receives the number of rows of e.stats_partial the kernel wrote (0: none)
static int run_fprop(const PPConvDesc& d, const Geo& geo, const void* x, const void* wf, const TapEpilogue& e,
                     bool* used_tc, void* col, void* wpad, cudaStream_t s,
                     int* stats_rows = nullptr) {
  TapGemm g;
  if (stats_rows) *stats_rows = 0;
  if (stem_direct_supported(d)) {
    if (used_tc) *used_tc = false;
    if (stats_rows && e.stats_partial) *stats_rows = stem_grid(d);
    return stem_fprop(d, x, wf, e, s);
  }
  const int Kpad = col_kpad(d, geo);
  if (Kpad && col && wpad) {
    PP_TRY(launch_im2col_small(d, x, col, geo.rows, geo.P, geo.Q, Kpad, s));
    PP_TRY(launch_pad_rows(wf, wpad, d.O, geo.T * d.C, Kpad, is_tf32(d) ? 1 : 0, s));
    plan_col(d, geo, Kpad, g);
    if (used_tc) *used_tc = true;
    if (stats_rows && e.stats_partial) *stats_rows = tapgemm_tcgen05_grid(g);
    return tapgemm_tcgen05(g, col, wpad, e, s);
  }
  plan_fprop(d, geo, g);
  const bool tc = use_tcgen05(d, tapgemm_tcgen05_supported(g));
  PP_REQUIRE(tc || d.algo != PP_ALGO_TCGEN05, PP_EUNSUPPORTED, "PP_ALGO_TCGEN05 requested but C=%d O=%d unsupported",
             d.C, d.O);
  if (used_tc) *used_tc = tc;
  if (tc) {
    if (stats_rows && e.stats_partial) *stats_rows = tapgemm_tcgen05_grid(g);
    return tapgemm_tcgen05(g, x, wf, e, s);
  }
  PP_REQUIRE(!is_tf32(d), PP_EUNSUPPORTED, "PP_DTYPE_TF32 conv needs C%%32==0 (or C*kh*kw <= 1024) and O%%64==0 (C=%d O=%d)",
             d.C, d.O);
  TapEpilogue e2 = e;
  e2.stats_partial = nullptr;
  return tapgemm_simt(g, x, wf, e2, s);
}

// dx_add (optional, bf16 [N,H,W,C]): added to the data gradient — in the kernel's epilogue where the gradient is one
// identity-mapped tensor-core launch (every stride-1 conv), by a separate pass otherwise.
static int run_dgrad(const PPConvDesc& d, const Geo& geo, const void* dz, const void* wd, void* dx, cudaStream_t s,
                     const void* dx_add = nullptr) {
  TapGemm phases[64];
  int nph = 0;
  bool any_empty = false;
  PP_REQUIRE(d.stride <= 8, PP_EUNSUPPORTED, "stride %d > 8", d.stride);
  for (int ph = 0; ph < d.stride; ++ph)
    for (int pw = 0; pw < d.stride; ++pw) {
      if (ph >= d.H || pw >= d.W) continue;
      if (plan_dgrad_phase(d, geo, ph, pw, phases[nph])) ++nph; else any_empty = true;
    }
  if (any_empty) PP_CHECK_CUDA(cudaMemsetAsync(dx, 0, (size_t)d.N * d.H * d.W * d.C * act_esz(d), s));
  PP_REQUIRE(!dx_add || !is_tf32(d), PP_EUNSUPPORTED, "dx_add is bf16-only");
  const bool fuse_add = dx_add && nph == 1 && !any_empty && phases[0].out_identity &&
                        use_tcgen05(d, tapgemm_tcgen05_supported(phases[0]));
  for (int i = 0; i < nph; ++i) {
    TapEpilogue e;
    e.out = dx; e.out_f32 = is_tf32(d) ? 1 : 0; e.scale = nullptr; e.shift = nullptr; e.relu = 0;
    e.stats_partial = nullptr;
    e.add_src = fuse_add ? dx_add : nullptr;
    const bool tc = use_tcgen05(d, tapgemm_tcgen05_supported(phases[i]));
    PP_REQUIRE(tc || !is_tf32(d), PP_EUNSUPPORTED, "PP_DTYPE_TF32 data gradient needs O%%32==0 and C%%64==0 (C=%d O=%d)",
               d.C, d.O);
    PP_REQUIRE(tc || d.algo != PP_ALGO_TCGEN05 || d.C % 64 != 0, PP_EUNSUPPORTED,
               "PP_ALGO_TCGEN05 requested but dgrad unsupported");
    if (tc) PP_TRY(tapgemm_tcgen05(phases[i], dz, wd, e, s));
    else PP_TRY(tapgemm_simt(phases[i], dz, wd, e, s));
  }
  if (dx_add && !fuse_add)
    PP_TRY(launch_add_inplace((__nv_bfloat16*)dx, (const __nv_bfloat16*)dx_add, (size_t)d.N * d.H * d.W * d.C, s));
  return PP_OK;
}

static int run_wgrad(const PPConvDesc& d, const Geo& geo, const void* dz, const void* x, float* dw, float* wpartial,
                     void* col, int splits, bool tc, cudaStream_t s, int accumulate = 0) {
  TapGemm g;
  if (stem_direct_supported(d)) {
    PP_TRY(stem_wgrad(d, x, dz, wpartial, s));
    return launch_wgrad_finalize(d, wpartial, splits, geo.T * d.C, dw, s, accumulate);
  }
  const int Kpad = col_kpad(d, geo);
  if (Kpad && col) {
    PP_TRY(launch_im2col_small(d, x, col, geo.rows, geo.P, geo.Q, Kpad, s));
    plan_col(d, geo, Kpad, g);
    PP_TRY(wgrad_tcgen05(g, col, dz, d.O, wpartial, splits, s));
    return launch_wgrad_finalize(d, wpartial, splits, Kpad, dw, s, accumulate);
  }
  plan_fprop(d, geo, g);
  PP_REQUIRE(tc || !is_tf32(d), PP_EUNSUPPORTED, "PP_DTYPE_TF32 weight gradient needs C%%32==0 and O%%64==0 (C=%d O=%d)",
             d.C, d.O);
  if (tc) PP_TRY(wgrad_tcgen05(g, x, dz, d.O, wpartial, splits, s));
  else PP_TRY(wgrad_simt(g, x, dz, d.O, wpartial, splits, s));
  return launch_wgrad_finalize(d, wpartial, splits, geo.T * d.C, dw, s, accumulate);
}

// The block as ONE cooperative kernel when every output tile stays resident in TMEM (igemm_sm100.cu:
// passport_fused_kernel).  Returns PP_OK with *done = true when it ran; *done = false means "not eligible, use the
// kernel sequence".  Exactly one of (gamma_in/beta_in) or (w_oihw, Ss, Sk -> gamma_out/beta_out) describes the affine.
static int try_fused_block(const PPConvDesc& d, const Geo& geo, const void* x, const void* wf, const FwdWs& ws,
                           FusedArgs a, cudaStream_t s, bool* done) {
  *done = false;
  if (d.norm != PP_NORM_BN_TRAIN || !a.z || !d.z_f32 || d.algo == PP_ALGO_SIMT || is_tf32(d)) return PP_OK;
  if (stem_direct_supported(d) || col_kpad(d, geo)) return PP_OK;
  TapGemm g;
  plan_fprop(d, geo, g);
  if (!passport_fused_supported(g)) return PP_OK;
  a.partial = ws.partial;
  a.barrier = ws.barrier;
  a.eps = d.eps; a.momentum = d.momentum; a.relu = d.relu;
  a.Cin = d.C; a.T = geo.T;
  PP_TRY(passport_fused_tcgen05(g, x, wf, a, s));
  *done = true;
  return PP_OK;
}

}  // namespace pp

using namespace pp;

// ================================================================== C ABI
extern "C" {

int pp_version(void) { return PP_ABI_VERSION; }
const char* pp_last_error(void) { return get_error(); }

int pp_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  DevInfo di;
  PP_TRY(query_device(&di));
  if (sm_count) *sm_count = di.sm;
  if (cc_major) *cc_major = di.cc_major;
  if (cc_minor) *cc_minor = di.cc_minor;
  return PP_OK;
}

int pp_workspace_bytes(const PPConvDesc* d, int which, size_t* bytes) {
  Geo geo;
  PP_TRY(geo_of(d, &geo));
  PP_REQUIRE(bytes != nullptr, PP_EBADARG, "bytes is NULL");
  if (which == PP_WS_FWD) {
    *bytes = carve_fwd(*d, geo, nullptr).total;
  } else if (which == PP_WS_BWD) {
    bool tc;
    const int splits = wgrad_splits_for(*d, geo, &tc);
    *bytes = carve_bwd(*d, geo, nullptr, splits).total;
  } else {
    set_error("unknown workspace kind %d", which);
    return PP_EBADARG;
  }
  return PP_OK;
}

int pp_weight_prep(const PPConvDesc* d, const float* w_oihw, void* w_fprop, void* w_dgrad, void* stream) {
  Geo geo;
  PP_TRY(geo_of(d, &geo));
  PP_TRY(check_device());
  PP_REQUIRE(w_oihw && w_fprop, PP_EBADARG, "weight pointers are NULL");
  return launch_weight_prep(*d, w_oihw, w_fprop, w_dgrad, (cudaStream_t)stream);
}

int pp_key_pool(const PPConvDesc* d, int Bk, const float* key_nchw, double* S, void* stream) {
  Geo geo;
  PP_TRY(geo_of(d, &geo));
  PP_TRY(check_device());
  PP_REQUIRE(key_nchw && S && Bk > 0, PP_EBADARG, "key pool: NULL pointer or empty key batch");
  return launch_key_pool(*d, Bk, key_nchw, S, (cudaStream_t)stream);
}

int pp_passport_affine_fwd(const PPConvDesc* d, const float* w_oihw, const double* S_skey, const double* S_key,
                           const float* b_sign, float alpha, float* gamma, float* beta, float* sign_loss,
                           float* sign_acc, void* stream) {
  Geo geo;
  PP_TRY(geo_of(d, &geo));
  PP_TRY(check_device());
  PP_REQUIRE(w_oihw && S_skey && S_key && gamma && beta, PP_EBADARG, "passport affine: NULL pointer");
  return launch_passport_affine_fwd(*d, w_oihw, S_skey, S_key, b_sign, alpha, gamma, beta,
                                    sign_loss, sign_acc, (cudaStream_t)stream);
}

int pp_passport_affine_bwd(const PPConvDesc* d, const double* S_skey, const double* S_key, const float* gamma,
                           const float* b_sign, float alpha, const float* g_gamma, const float* g_beta,
                           const float* g_loss, float* dw_oihw, int accumulate, void* stream) {
  Geo geo;
  PP_TRY(geo_of(d, &geo));
  PP_TRY(check_device());
  PP_REQUIRE(S_skey && S_key && dw_oihw, PP_EBADARG, "passport affine bwd: NULL pointer");
  PP_REQUIRE(!(g_loss && b_sign) || gamma, PP_EBADARG, "passport affine bwd: gamma needed for the sign-loss term");
  return launch_passport_affine_bwd(*d, S_skey, S_key, gamma, b_sign, alpha, g_gamma, g_beta, g_loss, dw_oihw,
                                    accumulate, (cudaStream_t)stream);
}

int pp_passport_key_grad(const PPConvDesc* d, int Bk, const float* w_oihw, const float* gamma, const float* b_sign,
                         float alpha, const float* g_gamma, const float* g_beta, const float* g_loss,
                         double* scratch, float* dskey_nchw, float* dkey_nchw, void* stream) {
  Geo geo;
  PP_TRY(geo_of(d, &geo));
  PP_TRY(check_device());
  PP_REQUIRE(w_oihw && scratch && Bk > 0, PP_EBADARG, "passport key grad: NULL pointer");
  PP_REQUIRE(!(g_loss && b_sign) || gamma, PP_EBADARG, "passport key grad: gamma needed for the sign-loss term");
  const int K = geo.T * d->C;
  return launch_passport_key_grad(*d, Bk, w_oihw, gamma, b_sign, alpha, g_gamma, g_beta, g_loss,
                                  scratch, scratch + K, dskey_nchw, dkey_nchw, (cudaStream_t)stream);
}

int pp_signature_verify(int nlayers, const PPSigLayer* layers, int32_t* matched, float* gamma_out, void* stream) {
  PP_TRY(check_device());
  PP_REQUIRE(nlayers >= 0 && nlayers <= PP_SIG_MAX_LAYERS, PP_EBADARG, "signature verify: %d layers (max %d per call)",
             nlayers, PP_SIG_MAX_LAYERS);
  if (nlayers == 0) return PP_OK;
  PP_REQUIRE(layers && matched, PP_EBADARG, "signature verify: NULL pointer");
  for (int i = 0; i < nlayers; ++i) {
    PP_REQUIRE(layers[i].w_oihw && layers[i].S_skey && layers[i].b_sign, PP_EBADARG,
               "signature verify: layer %d has a NULL pointer", i);
    PP_REQUIRE(layers[i].O > 0 && layers[i].K > 0 && layers[i].gamma_offset >= 0 && layers[i].C > 0 &&
                   layers[i].K % layers[i].C == 0,
               PP_EBADSHAPE, "signature verify: layer %d has O=%d K=%d C=%d", i, layers[i].O, layers[i].K,
               layers[i].C);
  }
  return launch_signature_verify(nlayers, layers, matched, gamma_out, (cudaStream_t)stream);
}

int pp_sign_loss_fwd(int O, const float* gamma, const float* b_sign, float alpha, float* sign_loss, float* sign_acc,
                     void* stream) {
  PP_TRY(check_device());
  PP_REQUIRE(O > 0 && gamma && b_sign, PP_EBADARG, "sign loss: bad arguments");
  return launch_sign_loss_fwd(O, gamma, b_sign, alpha, sign_loss, sign_acc, (cudaStream_t)stream);
}

int pp_sign_loss_bwd(int O, const float* gamma, const float* b_sign, float alpha, const float* g_loss,
                     float* g_gamma, void* stream) {
  PP_TRY(check_device());
  PP_REQUIRE(O > 0 && gamma && b_sign && g_gamma, PP_EBADARG, "sign loss bwd: bad arguments");
  return launch_sign_loss_bwd(O, gamma, b_sign, alpha, g_loss, g_gamma, (cudaStream_t)stream);
}

int pp_conv_block_fwd(const PPConvDesc* d, const void* x, const void* w_fprop, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, void* z, void* y, float* save_mean,
                      float* save_invstd, void* workspace, size_t ws_bytes, void* stream) {
  return pp_conv_block_fwd_res(d, x, w_fprop, gamma, beta, running_mean, running_var, z, y, save_mean, save_invstd,
                               nullptr, workspace, ws_bytes, stream);
}

int pp_conv_block_fwd_res(const PPConvDesc* d, const void* x, const void* w_fprop, const float* gamma,
                          const float* beta, float* running_mean, float* running_var, void* z, void* y,
                          float* save_mean, float* save_invstd, const void* residual, void* workspace, size_t ws_bytes,
                          void* stream) {
  Geo geo;
  PP_TRY(geo_of(d, &geo));
  PP_TRY(check_device());
  cudaStream_t s = (cudaStream_t)stream;
  PP_REQUIRE(x && w_fprop && y, PP_EBADARG, "conv block fwd: NULL pointer");
  PP_REQUIRE(d->norm == PP_NORM_NONE || d->norm == PP_NORM_BN_TRAIN || d->norm == PP_NORM_BN_EVAL ||
                 d->norm == PP_NORM_GN,
             PP_EBADARG, "unknown norm %d", d->norm);
  PP_REQUIRE(d->norm != PP_NORM_GN || z, PP_EBADARG, "group / instance norm needs the z buffer");
  PP_REQUIRE(d->norm != PP_NORM_GN || (save_mean && save_invstd), PP_EBADARG,
             "group / instance norm needs save_mean / save_invstd [N*groups]");
  PP_REQUIRE(d->norm != PP_NORM_BN_EVAL || (running_mean && running_var), PP_EBADARG, "BN eval needs running stats");
  PP_REQUIRE(d->norm != PP_NORM_BN_TRAIN || z, PP_EBADARG, "BN train needs the z buffer");
  PP_REQUIRE(!is_tf32(*d) || !z || d->z_f32, PP_EBADARG, "PP_DTYPE_TF32 keeps z in fp32 (set z_f32)");
  PP_REQUIRE(!residual || (z && !is_tf32(*d) && d->norm != PP_NORM_GN), PP_EUNSUPPORTED,
             "the fused residual join needs the z buffer, bf16 tensors and a batch-norm / plain block");
  const int af32 = is_tf32(*d) ? 1 : 0;
  FwdWs ws = carve_fwd(*d, geo, workspace);
  PP_REQUIRE(workspace && ws_bytes >= ws.total, PP_EWORKSPACE, "fwd workspace too small: need %zu, got %zu", ws.total,
             ws_bytes);
  if (z == nullptr) {
    // inference: statistics are known up front, the affine lives in the conv epilogue
    PP_TRY(launch_bn_finalize(*d, (int)geo.rows, nullptr, 0, gamma, beta, running_mean, running_var, save_mean,
                              save_invstd, ws.ca, ws.cb, s));
    TapEpilogue e;
    e.out = y; e.out_f32 = af32; e.scale = ws.ca; e.shift = ws.cb; e.relu = d->relu; e.stats_partial = nullptr;
    return run_fprop(*d, geo, x, w_fprop, e, nullptr, ws.col, ws.wpad, s);
  }
  {
    FusedArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.y = y; fa.z = d->z_f32 ? (float*)z : nullptr; fa.gamma_in = gamma; fa.beta_in = beta;
    fa.rmean = running_mean; fa.rvar = running_var; fa.save_mean = save_mean; fa.save_invstd = save_invstd;
    bool done = false;
    if (!residual) PP_TRY(try_fused_block(*d, geo, x, w_fprop, ws, fa, s, &done));
    if (done) return PP_OK;
  }
  TapEpilogue e;
  e.out = z; e.out_f32 = d->z_f32; e.scale = nullptr; e.shift = nullptr; e.relu = 0;
  const bool fused_stats = d->norm == PP_NORM_BN_TRAIN && d->O <= tapgemm_tcgen05_max_stats_width();
  e.stats_partial = fused_stats ? ws.partial : nullptr;
  bool tc = false;
  int stats_rows = 0;   // rows of ws.partial the conv kernel's epilogue wrote (0: it has no fused statistics)
  PP_TRY(run_fprop(*d, geo, x, w_fprop, e, &tc, ws.col, ws.wpad, s, &stats_rows));
  if (d->norm == PP_NORM_GN)
    return launch_gn_fwd(*d, geo.P * geo.Q, z, gamma, beta, save_mean, save_invstd, ws.ca, ws.cb, ws.partial,
                         (__nv_bfloat16*)y, s);
  int num_partials = 0;
  if (d->norm == PP_NORM_BN_TRAIN) {
    if (stats_rows > 0) {
      num_partials = stats_rows;
    } else {
      PP_TRY(launch_col_stats(z, d->z_f32, geo.rows, d->O, ws.partial, &num_partials, s));
    }
  }
  PP_TRY(launch_bn_finalize(*d, (int)geo.rows, ws.partial, num_partials, gamma, beta, running_mean, running_var,
                            save_mean, save_invstd, ws.ca, ws.cb, s));
  // algorithmic bytes of the pass: read z once, write y (bf16) once
  const double zb = d->z_f32 ? 4.0 : 2.0;
  prof_begin(PROF_AFFINE, (double)geo.rows * d->O * (zb + (double)act_esz(*d) + (residual ? 2.0 : 0.0)), d->C, d->O,
             geo.T, s);
  const int rc = launch_affine_apply(z, d->z_f32, geo.rows, d->O, ws.ca, ws.cb, d->relu, y, af32, residual, s);
  prof_end(PROF_AFFINE, s);
  return rc;
}

}  // extern "C"

// dz_ext != NULL: dz goes to that caller-owned buffer instead of the workspace and the weight gradient is NOT computed
// (the caller runs pp_conv_wgrad on it, possibly on another stream: pp_conv_block_bwd_dz).
static int conv_block_bwd_impl(const PPConvDesc* d, const void* dy, const void* x, const void* w_dgrad, const void* z,
                               const float* gamma, const float* beta, const float* save_mean,
                               const float* save_invstd, void* dx, float* dw_oihw, float* dgamma, float* dbeta,
                               void* dz_ext, const void* dx_add, void* workspace, size_t ws_bytes, void* stream) {
  Geo geo;
  PP_TRY(geo_of(d, &geo));
  PP_TRY(check_device());
  cudaStream_t s = (cudaStream_t)stream;
  PP_REQUIRE(dy && z && save_mean && save_invstd && dgamma && dbeta, PP_EBADARG, "conv block bwd: NULL pointer");
  PP_REQUIRE(!dx || w_dgrad, PP_EBADARG, "dx requested without w_dgrad");
  PP_REQUIRE(!dw_oihw || x, PP_EBADARG, "dw requested without x");
  PP_REQUIRE(!is_tf32(*d) || d->z_f32, PP_EBADARG, "PP_DTYPE_TF32 keeps z in fp32 (set z_f32)");
  const int af32 = is_tf32(*d) ? 1 : 0;
  const double ab = (double)act_esz(*d);
  bool wg_tc = false;
  const int splits = wgrad_splits_for(*d, geo, &wg_tc);
  BwdWs ws = carve_bwd(*d, geo, workspace, splits);
  PP_REQUIRE(workspace && ws_bytes >= ws.total, PP_EWORKSPACE, "bwd workspace too small: need %zu, got %zu", ws.total,
             ws_bytes);
  PP_REQUIRE(!dz_ext || d->norm != PP_NORM_GN, PP_EUNSUPPORTED, "the split backward is not built for group / instance norm");
  if (dz_ext) ws.dz = dz_ext;
  if (d->norm == PP_NORM_GN) {
    const int HW = geo.P * geo.Q;
    PP_TRY(launch_gn_bwd_reduce(*d, HW, (const __nv_bfloat16*)dy, z, gamma, beta, save_mean, save_invstd, ws.ca, ws.cb,
                                ws.k1, ws.k2, ws.k3, ws.partial, ws.contrib, dgamma, dbeta, s));
    if (!dx && !dw_oihw) return PP_OK;
    PP_TRY(launch_gn_dz(*d, HW, (const __nv_bfloat16*)dy, z, ws.ca, ws.cb, ws.k1, ws.k2, ws.k3, (__nv_bfloat16*)ws.dz,
                        s));
    if (dx) PP_TRY(run_dgrad(*d, geo, ws.dz, w_dgrad, dx, s));
    if (dw_oihw)
      PP_TRY(run_wgrad(*d, geo, ws.dz, x, dw_oihw, ws.wpartial, ws.col, splits, wg_tc, s, d->flags & PP_FLAG_ACC_DW));
    return PP_OK;
  }
  int num_partials = 0;
  const double zb = d->z_f32 ? 4.0 : 2.0;
  prof_begin(PROF_REDUCE, (double)geo.rows * d->O * (zb + ab), d->C, d->O, geo.T, s);   // read dy + z
  const int share = (d->flags & PP_FLAG_SHARE_SM) ? 1 : 0;
  const int rc_red = launch_bwd_reduce(dy, af32, z, d->z_f32, geo.rows, d->O, gamma, beta, save_mean, save_invstd,
                                       d->relu, ws.partial, &num_partials, s, share);
  prof_end(PROF_REDUCE, s);
  PP_TRY(rc_red);
  PP_TRY(launch_bwd_coef(*d, geo.rows, ws.partial, num_partials, gamma, beta, save_mean, save_invstd, dgamma, dbeta,
                         ws.k1, ws.k2, ws.k3, ws.ca, ws.cb, s));
  if (!dx && !dw_oihw && !dz_ext) return PP_OK;
  prof_begin(PROF_DZ, (double)geo.rows * d->O * (zb + 2.0 * ab), d->C, d->O, geo.T, s);   // read dy + z, write dz
  const int rc_dz = launch_bwd_dz(dy, af32, z, d->z_f32, geo.rows, d->O, ws.ca, ws.cb, d->relu, ws.k1, ws.k2, ws.k3,
                                  ws.dz, s, share);
  prof_end(PROF_DZ, s);
  PP_TRY(rc_dz);
  if (dx) PP_TRY(run_dgrad(*d, geo, ws.dz, w_dgrad, dx, s, dx_add));
  if (dw_oihw && !dz_ext)
    PP_TRY(run_wgrad(*d, geo, ws.dz, x, dw_oihw, ws.wpartial, ws.col, splits, wg_tc, s, d->flags & PP_FLAG_ACC_DW));
  return PP_OK;
}

extern "C" {

int pp_conv_block_bwd(const PPConvDesc* d, const void* dy, const void* x, const void* w_dgrad, const void* z,
                      const float* gamma, const float* beta, const float* save_mean, const float* save_invstd,
                      void* dx, float* dw_oihw, float* dgamma, float* dbeta, void* workspace, size_t ws_bytes,
                      void* stream) {
  return conv_block_bwd_impl(d, dy, x, w_dgrad, z, gamma, beta, save_mean, save_invstd, dx, dw_oihw, dgamma, dbeta,
                             nullptr, nullptr, workspace, ws_bytes, stream);
}

int pp_conv_block_bwd_dz(const PPConvDesc* d, const void* dy, const void* w_dgrad, const void* z, const float* gamma,
                         const float* beta, const float* save_mean, const float* save_invstd, void* dx,
                         const void* dx_add, float* dgamma, float* dbeta, void* dz_out, void* workspace,
                         size_t ws_bytes, void* stream) {
  PP_REQUIRE(dz_out != nullptr, PP_EBADARG, "conv block bwd (split): dz_out is NULL");
  PP_REQUIRE(!dx_add || dx, PP_EBADARG, "conv block bwd (split): dx_add without dx");
  return conv_block_bwd_impl(d, dy, nullptr, w_dgrad, z, gamma, beta, save_mean, save_invstd, dx, nullptr, dgamma,
                             dbeta, dz_out, dx_add, workspace, ws_bytes, stream);
}

int pp_passport_conv_fwd(const PPConvDesc* d, const void* x, const void* w_fprop, const float* w_oihw,
                         const double* S_skey, const double* S_key, const float* scale_pub, const float* bias_pub,
                         const float* b_sign, float alpha, float* running_mean, float* running_var, void* y, void* z,
                         float* gamma, float* beta, float* save_mean, float* save_invstd, float* sign_loss,
                         float* sign_acc, void* workspace, size_t ws_bytes, void* stream) {
  Geo geo;
  PP_TRY(geo_of(d, &geo));
  PP_TRY(check_device());
  cudaStream_t s = (cudaStream_t)stream;
  const bool pub = scale_pub != nullptr && bias_pub != nullptr;
  PP_REQUIRE(x && w_fprop && y, PP_EBADARG, "passport conv fwd: NULL pointer");
  PP_REQUIRE(pub || (w_oihw && S_skey && S_key && gamma && beta), PP_EBADARG,
             "passport conv fwd: the passport path needs w_oihw, S_skey, S_key and the gamma / beta buffers");
  PP_REQUIRE((scale_pub == nullptr) == (bias_pub == nullptr), PP_EBADARG,
             "passport conv fwd: scale_pub and bias_pub must both be given or both be NULL");
  FwdWs ws = carve_fwd(*d, geo, workspace);
  PP_REQUIRE(workspace && ws_bytes >= ws.total, PP_EWORKSPACE, "fwd workspace too small: need %zu, got %zu", ws.total,
             ws_bytes);
  {
    FusedArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.y = y; fa.z = (z && d->z_f32) ? (float*)z : nullptr;
    if (pub) { fa.gamma_in = scale_pub; fa.beta_in = bias_pub; }
    else {
      fa.w_oihw = w_oihw; fa.Ss = S_skey; fa.Sk = S_key; fa.gamma_out = gamma; fa.beta_out = beta;
      fa.b_sign = b_sign; fa.alpha = alpha; fa.sign_loss = sign_loss; fa.sign_acc = sign_acc;
    }
    fa.rmean = running_mean; fa.rvar = running_var; fa.save_mean = save_mean; fa.save_invstd = save_invstd;
    bool done = false;
    PP_TRY(try_fused_block(*d, geo, x, w_fprop, ws, fa, s, &done));
    if (done) return PP_OK;
  }
  // kernel sequence: gamma / beta (+ sign loss) first, then the conv block
  if (!pub)
    PP_TRY(launch_passport_affine_fwd(*d, w_oihw, S_skey, S_key, b_sign, alpha, gamma, beta, sign_loss, sign_acc, s));
  return pp_conv_block_fwd(d, x, w_fprop, pub ? scale_pub : gamma, pub ? bias_pub : beta, running_mean, running_var,
                           z, y, save_mean, save_invstd, workspace, ws_bytes, stream);
}

int pp_passport_conv_bwd(const PPConvDesc* d, const void* dy, const void* x, const void* w_dgrad, const void* z,
                         const float* gamma, const float* beta, const float* save_mean, const float* save_invstd,
                         const double* S_skey, const double* S_key, const float* b_sign, float alpha,
                         const float* g_sign_loss, void* dx, float* dw_oihw, float* dgamma, float* dbeta,
                         void* workspace, size_t ws_bytes, void* stream) {
  PP_REQUIRE(d != nullptr, PP_EBADARG, "PPConvDesc is NULL");
  PP_REQUIRE((S_skey == nullptr) == (S_key == nullptr), PP_EBADARG, "passport conv bwd: S_skey / S_key mismatch");
  PP_REQUIRE(!S_skey || dw_oihw, PP_EBADARG, "passport conv bwd: the passport path needs dw_oihw");
  PPConvDesc dd = *d;
  dd.flags &= ~(PP_FLAG_ACC_DGAMMA | PP_FLAG_ACC_DBETA);     // dgamma / dbeta are consumed below, never accumulated
  PP_TRY(pp_conv_block_bwd(&dd, dy, x, w_dgrad, z, gamma, beta, save_mean, save_invstd, dx, dw_oihw, dgamma, dbeta,
                           workspace, ws_bytes, stream));
  if (!S_skey) return PP_OK;                                  // public path: dgamma / dbeta ARE dscale / dbias
  // rank-1 term of the weight gradient: gamma, beta are functions of W (passportconv2d.py:146-152, 167-173)
  return launch_passport_affine_bwd(*d, S_skey, S_key, gamma, b_sign, alpha, dgamma, dbeta, g_sign_loss, dw_oihw,
                                    /*accumulate=*/1, (cudaStream_t)stream);
}

int pp_conv_fwd_raw(const PPConvDesc* d, const void* x, const void* w_fprop, void* z, void* workspace,
                    size_t ws_bytes, void* stream) {
  Geo geo;
  PP_TRY(geo_of(d, &geo));
  PP_TRY(check_device());
  PP_REQUIRE(x && w_fprop && z, PP_EBADARG, "conv fwd raw: NULL pointer");
  FwdWs ws = carve_fwd(*d, geo, workspace);
  const bool have_ws = workspace != nullptr && ws_bytes >= ws.total;
  TapEpilogue e;
  e.out = z; e.out_f32 = d->z_f32; e.scale = nullptr; e.shift = nullptr; e.relu = 0; e.stats_partial = nullptr;
  return run_fprop(*d, geo, x, w_fprop, e, nullptr, have_ws ? ws.col : nullptr, have_ws ? ws.wpad : nullptr,
                   (cudaStream_t)stream);
}

int pp_conv_dgrad(const PPConvDesc* d, const void* dz, const void* w_dgrad, void* dx, void* stream) {
  Geo geo;
  PP_TRY(geo_of(d, &geo));
  PP_TRY(check_device());
  PP_REQUIRE(dz && w_dgrad && dx, PP_EBADARG, "conv dgrad: NULL pointer");
  return run_dgrad(*d, geo, dz, w_dgrad, dx, (cudaStream_t)stream);
}

int pp_conv_wgrad(const PPConvDesc* d, const void* dz, const void* x, float* dw_oihw, void* workspace,
                  size_t ws_bytes, void* stream) {
  Geo geo;
  PP_TRY(geo_of(d, &geo));
  PP_TRY(check_device());
  PP_REQUIRE(dz && x && dw_oihw, PP_EBADARG, "conv wgrad: NULL pointer");
  bool tc = false;
  const int splits = wgrad_splits_for(*d, geo, &tc);
  BwdWs ws = carve_bwd(*d, geo, workspace, splits);
  PP_REQUIRE(workspace && ws_bytes >= ws.total, PP_EWORKSPACE, "wgrad workspace too small: need %zu, got %zu",
             ws.total, ws_bytes);
  return run_wgrad(*d, geo, dz, x, dw_oihw, ws.wpartial, ws.col, splits, tc, (cudaStream_t)stream,
                   d->flags & PP_FLAG_ACC_DW);
}

int pp_sgd_step(size_t n, float* param, const float* grad, float* momentum_buf, float lr, float momentum,
                float weight_decay, int first_step, void* stream) {
  PP_TRY(check_device());
  PP_REQUIRE(param && grad && (momentum == 0.0f || momentum_buf), PP_EBADARG, "sgd: NULL pointer");
  if (n == 0) return PP_OK;
  return launch_sgd(n, param, grad, momentum_buf, lr, momentum, weight_decay, first_step, (cudaStream_t)stream);
}

int pp_sgd_step_dev(size_t n, float* param, const float* grad, float* momentum_buf, const float* hyper,
                    void* stream) {
  PP_TRY(check_device());
  PP_REQUIRE(param && grad && momentum_buf && hyper, PP_EBADARG, "sgd (device hyper-parameters): NULL pointer");
  if (n == 0) return PP_OK;
  return launch_sgd_dev(n, param, grad, momentum_buf, hyper, (cudaStream_t)stream);
}

int pp_ce_top1(int N, int classes, const void* logits, int logits_bf16, const int64_t* target, float* loss,
               float* top1, float* dlogits, int accumulate, void* stream) {
  PP_TRY(check_device());
  PP_REQUIRE(N > 0 && classes > 0, PP_EBADSHAPE, "ce_top1: N=%d classes=%d", N, classes);
  PP_REQUIRE(logits && target && (loss || top1 || dlogits), PP_EBADARG, "ce_top1: NULL pointer");
  return launch_ce_top1(N, classes, logits, logits_bf16, (const long long*)target, loss, top1, dlogits, accumulate,
                        (cudaStream_t)stream);
}

int pp_add_relu_fwd(size_t n, const void* a, const void* b, void* y, void* stream) {
  PP_TRY(check_device());
  PP_REQUIRE(a && b && y, PP_EBADARG, "add_relu: NULL pointer");
  if (n == 0) return PP_OK;
  return launch_add_relu_fwd((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (__nv_bfloat16*)y, n,
                             (cudaStream_t)stream);
}

int pp_add_relu_bwd(size_t n, const void* gy, const void* y, void* gx, void* stream) {
  PP_TRY(check_device());
  PP_REQUIRE(gy && y && gx, PP_EBADARG, "add_relu bwd: NULL pointer");
  if (n == 0) return PP_OK;
  return launch_add_relu_bwd((const __nv_bfloat16*)gy, (const __nv_bfloat16*)y, (__nv_bfloat16*)gx, n,
                             (cudaStream_t)stream);
}

int pp_maxpool_fwd(int N, int H, int W, int C, int k, int stride, int pad, const void* x, int f32, void* y,
                   uint8_t* argmax, void* stream) {
  PP_TRY(check_device());
  PP_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, PP_EBADSHAPE, "maxpool: N=%d H=%d W=%d C=%d (C%%8)", N, H, W, C);
  PP_REQUIRE(k >= 1 && k <= 15 && stride >= 1 && pad >= 0 && 2 * pad <= k && H + 2 * pad >= k && W + 2 * pad >= k,
             PP_EBADSHAPE, "maxpool: k=%d stride=%d pad=%d", k, stride, pad);
  PP_REQUIRE(x && y && argmax, PP_EBADARG, "maxpool: NULL pointer");
  return launch_maxpool_fwd(N, H, W, C, k, stride, pad, x, f32, y, argmax, (cudaStream_t)stream);
}

int pp_maxpool_bwd(int N, int H, int W, int C, int k, int stride, int pad, const void* dy, const uint8_t* argmax,
                   int f32, void* dx, void* stream) {
  PP_TRY(check_device());
  PP_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, PP_EBADSHAPE, "maxpool: N=%d H=%d W=%d C=%d (C%%8)", N, H, W, C);
  PP_REQUIRE(k >= 1 && k <= 15 && stride >= 1 && pad >= 0 && 2 * pad <= k && H + 2 * pad >= k && W + 2 * pad >= k,
             PP_EBADSHAPE, "maxpool: k=%d stride=%d pad=%d", k, stride, pad);
  PP_REQUIRE(dy && dx && argmax, PP_EBADARG, "maxpool bwd: NULL pointer");
  return launch_maxpool_bwd(N, H, W, C, k, stride, pad, dy, argmax, f32, dx, (cudaStream_t)stream);
}

int pp_debug_last_timeout(void) { return debug_last_timeout(); }
int pp_debug_fused(int on) { return debug_fused(on); }

long long pp_launch_count(int reset) {
  const long long v = g_launches.load();
  if (reset) g_launches.store(0);
  return v;
}

int pp_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
  if (g_prof_on) {  // a new recording session starts empty
    for (int k = 0; k < PROF_KINDS; ++k) {
      for (auto& r : g_prof[k]) { cudaEventDestroy(r.start); cudaEventDestroy(r.stop); }
      g_prof[k].clear();
    }
  }
  return PP_OK;
}

int pp_profile_read(int kind, int filter_c, int filter_nout, int filter_taps, double* total_ms,
                    double* total_flops, int* launches) {
  PP_REQUIRE(kind >= 0 && kind < PROF_KINDS, PP_EBADARG, "unknown profile kind %d", kind);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double ms = 0.0, fl = 0.0;
  int n = 0;
  for (auto& r : g_prof[kind]) {
    float t = 0.f;
    const bool match = (filter_c <= 0 || r.c == filter_c) && (filter_nout <= 0 || r.nout == filter_nout) &&
                       (filter_taps <= 0 || r.taps == filter_taps);
    if (match && cudaEventSynchronize(r.stop) == cudaSuccess &&
        cudaEventElapsedTime(&t, r.start, r.stop) == cudaSuccess) {
      ms += t; fl += r.flops; ++n;
    }
  }
  cudaGetLastError();
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = fl;
  if (launches) *launches = n;
  return PP_OK;
}

}  // extern "C"
