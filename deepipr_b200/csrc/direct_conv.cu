// direct_conv.cu — SIMT tap-GEMM and weight gradient for shapes the tensor-core kernels do not take
// (the 3-channel network stem, channel counts that are not multiples of 64).  Same TapGemm contract as
// igemm_sm100.cu, so the two can be compared tile for tile in the tests.
#include "common.h"

namespace pp {

struct SimtDev {
  int M, N, H, W, C, P, Q, PQ;
  int base_h, base_w, step_h, step_w;
  int ntaps, Nout, Ktot;
  int out_H, out_W, out_sh, out_sw, out_ph, out_pw, out_identity;
  int8_t tap_dh[kMaxTaps], tap_dw[kMaxTaps];
  int tap_kofs[kMaxTaps];
  void* out;
  int out_f32;
  const float* scale;
  const float* shift;
  int relu;
};

static void fill_simt(SimtDev& p, const TapGemm& g) {
  p.M = g.N * g.P * g.Q; p.N = g.N; p.H = g.H; p.W = g.W; p.C = g.C; p.P = g.P; p.Q = g.Q; p.PQ = g.P * g.Q;
  p.base_h = g.base_h; p.base_w = g.base_w; p.step_h = g.step_h; p.step_w = g.step_w;
  p.ntaps = g.ntaps; p.Nout = g.Nout; p.Ktot = g.Ktot;
  p.out_H = g.out_H; p.out_W = g.out_W; p.out_sh = g.out_sh; p.out_sw = g.out_sw; p.out_ph = g.out_ph;
  p.out_pw = g.out_pw; p.out_identity = g.out_identity;
  for (int t = 0; t < g.ntaps; ++t) { p.tap_dh[t] = g.tap_dh[t]; p.tap_dw[t] = g.tap_dw[t]; p.tap_kofs[t] = g.tap_kofs[t]; }
}

// one thread per output element (n fastest => activation loads are warp-broadcast, stores coalesced);
// the weight matrix is staged transposed in shared memory when it is small (the stem: 64 x 27).
template <bool W_IN_SMEM>
__global__ void tapgemm_simt_kernel(const __grid_constant__ SimtDev p, const __nv_bfloat16* __restrict__ act,
                                    const __nv_bfloat16* __restrict__ B) {
  extern __shared__ float s_w[];  // [Ktot][Nout] when W_IN_SMEM
  if (W_IN_SMEM) {
    for (int i = threadIdx.x; i < p.Nout * p.Ktot; i += blockDim.x) {
      const int n = i / p.Ktot, k = i % p.Ktot;
      s_w[k * p.Nout + n] = __bfloat162float(B[i]);
    }
    __syncthreads();
  }
  const size_t total = (size_t)p.M * p.Nout;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx % p.Nout);
    const int m = (int)(idx / p.Nout);
    const int img = m / p.PQ;
    const int rem = m - img * p.PQ;
    const int pp_ = rem / p.Q;
    const int qq_ = rem - pp_ * p.Q;
    float acc = 0.0f;
    for (int t = 0; t < p.ntaps; ++t) {
      const int h = p.base_h + pp_ * p.step_h + p.tap_dh[t];
      const int w = p.base_w + qq_ * p.step_w + p.tap_dw[t];
      if (h < 0 || h >= p.H || w < 0 || w >= p.W) continue;
      const __nv_bfloat16* a = act + (((size_t)img * p.H + h) * p.W + w) * p.C;
      const int kofs = p.tap_kofs[t];
      if (W_IN_SMEM) {
        for (int c = 0; c < p.C; ++c) acc = fmaf(__bfloat162float(a[c]), s_w[(kofs + c) * p.Nout + n], acc);
      } else {
        const __nv_bfloat16* b = B + (size_t)n * p.Ktot + kofs;
        for (int c = 0; c < p.C; ++c) acc = fmaf(__bfloat162float(a[c]), __bfloat162float(b[c]), acc);
      }
    }
    if (p.scale) acc *= p.scale[n];
    if (p.shift) acc += p.shift[n];
    if (p.relu) acc = fmaxf(acc, 0.0f);
    size_t out_row = (size_t)m;
    if (!p.out_identity)
      out_row = ((size_t)img * p.out_H + (size_t)(pp_ * p.out_sh + p.out_ph)) * p.out_W + (size_t)(qq_ * p.out_sw + p.out_pw);
    if (p.out_f32) reinterpret_cast<float*>(p.out)[out_row * p.Nout + n] = acc;
    else reinterpret_cast<__nv_bfloat16*>(p.out)[out_row * p.Nout + n] = __float2bfloat16_rn(acc);
  }
}

int tapgemm_simt(const TapGemm& g, const void* act, const void* B, const TapEpilogue& e, cudaStream_t s) {
  PP_REQUIRE(e.stats_partial == nullptr, PP_EUNSUPPORTED, "SIMT tap-GEMM has no fused statistics");
  SimtDev p;
  fill_simt(p, g);
  p.out = e.out; p.out_f32 = e.out_f32; p.scale = e.scale; p.shift = e.shift; p.relu = e.relu;
  const size_t total = (size_t)p.M * p.Nout;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  const size_t wbytes = (size_t)g.Nout * g.Ktot * sizeof(float);
  if (wbytes <= 40 * 1024) {
    tapgemm_simt_kernel<true><<<(int)blocks, 256, wbytes, s>>>(p, (const __nv_bfloat16*)act, (const __nv_bfloat16*)B);
  } else {
    tapgemm_simt_kernel<false><<<(int)blocks, 256, 0, s>>>(p, (const __nv_bfloat16*)act, (const __nv_bfloat16*)B);
  }
  PP_POST_LAUNCH();
  return PP_OK;
}

// partial[split][o][t*C + c] = sum over the split's pixels of dz[m, o] * x[pixel(m) + tap t, c]
// one thread per (o, k) with o fastest (dz loads coalesced, x loads warp-broadcast).
__global__ void wgrad_simt_kernel(const __grid_constant__ SimtDev p, const __nv_bfloat16* __restrict__ x,
                                  const __nv_bfloat16* __restrict__ dz, int O, float* __restrict__ partial,
                                  int pixels_per_split) {
  const int split = blockIdx.y;
  const int m_lo = split * pixels_per_split;
  int m_hi = m_lo + pixels_per_split;
  if (m_hi > p.M) m_hi = p.M;
  const int total = O * p.Ktot;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int o = idx % O;
    const int k = idx / O;
    const int t = k / p.C;
    const int c = k - t * p.C;
    const int dh = p.tap_dh[t], dw = p.tap_dw[t];
    float acc = 0.0f;
    for (int m = m_lo; m < m_hi; ++m) {
      const int img = m / p.PQ;
      const int rem = m - img * p.PQ;
      const int pp_ = rem / p.Q;
      const int qq_ = rem - pp_ * p.Q;
      const int h = p.base_h + pp_ * p.step_h + dh;
      const int w = p.base_w + qq_ * p.step_w + dw;
      if (h < 0 || h >= p.H || w < 0 || w >= p.W) continue;
      const float xv = __bfloat162float(x[(((size_t)img * p.H + h) * p.W + w) * p.C + c]);
      acc = fmaf(__bfloat162float(dz[(size_t)m * O + o]), xv, acc);
    }
    partial[((size_t)split * O + o) * p.Ktot + k] = acc;
  }
}

int wgrad_simt_pick_splits(const TapGemm& g, int O) {
  const long long M = (long long)g.N * g.P * g.Q;
  long long splits = (M + 1023) / 1024;
  if (splits > 2048) splits = 2048;
  const long long per_split = (long long)O * g.ntaps * g.C * 4;
  while (splits > 1 && per_split * splits > (64ll << 20)) --splits;
  if (splits < 1) splits = 1;
  return (int)splits;
}

int wgrad_simt(const TapGemm& g, const void* x, const void* dz, int O, float* partial, int splits, cudaStream_t s) {
  SimtDev p;
  fill_simt(p, g);
  p.out = nullptr; p.out_f32 = 0; p.scale = nullptr; p.shift = nullptr; p.relu = 0;
  p.Ktot = g.ntaps * g.C;
  const int total = O * p.Ktot;
  const int pixels_per_split = (p.M + splits - 1) / splits;
  int bx = (total + 127) / 128;
  if (bx > 4096) bx = 4096;
  dim3 grid(bx, splits);
  wgrad_simt_kernel<<<grid, 128, 0, s>>>(p, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dz, O, partial,
                                        pixels_per_split);
  PP_POST_LAUNCH();
  return PP_OK;
}

}  // namespace pp
