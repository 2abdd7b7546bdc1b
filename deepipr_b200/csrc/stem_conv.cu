// stem_conv.cu — the 3-channel network stem (ConvBlock(3, 64, 3, 1, 1), models/resnet_normal.py:66,
// resnet_passport_private.py:100) as direct CUDA-core kernels.
//
// With C = 3 the contraction is K = 27: 0.15 % of the network's FLOPs, but as an im2col + tensor-core GEMM it moved
// 4x the bytes of its output (a [pixels, 64] bf16 im2col matrix written and re-read) and cost 6 % of the training
// step.  Here the layer is bound by its output write instead:
//   fprop : lane = 2 output channels (their 2 x 27 weights live in registers), a warp walks one output row in strips
//           of 4 pixels whose 3 x 6 x 3 input patch is read from a zero-padded shared-memory halo tile with
//           broadcast 128-bit loads; every pixel is one coalesced 256-byte (fp32) / 128-byte (bf16) store per warp;
//           the BatchNorm statistics (sum z, sum z^2 per channel) are plain per-lane running sums.
//   wgrad : same walk, 2 x 27 accumulators per lane, the tile's dz staged in shared memory by 128-bit loads; warps
//           are combined in a fixed order in shared memory, blocks by wgrad_finalize (deterministic).
// The data gradient of the stem is not needed in training (images carry no gradient); when it is requested the
// generic tap-GEMM path computes it.
#include <stdlib.h>

#include "common.h"

namespace pp {

constexpr int kStemThreads = 256;  // 8 warps = 8 output rows of a tile
constexpr int kStemTH = 8;         // tile: 8 output rows x 32 output columns
constexpr int kStemTW = 32;
constexpr int kStemHaloRows = kStemTH + 2;
constexpr int kStemRowF = 104;     // halo row: 34 columns x 3 channels = 102 floats, padded to a 16-byte multiple
constexpr int kStemK = 27;
constexpr int kStemO = 64;

struct StemDev {
  int N, H, W;               // input == output spatial size (3x3, stride 1, pad 1)
  int tiles_h, tiles_w, num_tiles;
  // fprop epilogue
  void* out;                 // [N*H*W, 64] fp32 or bf16
  int out_f32;
  const float* scale;        // per-channel affine (NULL => 1 / 0)
  const float* shift;
  int relu;
  float* stats_partial;      // NULL or [gridDim.x][2][64]
};

bool stem_direct_supported(const PPConvDesc& d) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("PP_STEM_DIRECT");   // =0 restores the im2col + tensor-core path (A/B comparison)
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  return enabled && d.algo != PP_ALGO_SIMT && d.dtype == PP_DTYPE_BF16 && d.C == 3 && d.O == kStemO && d.kh == 3 && d.kw == 3 && d.stride == 1 &&
         d.pad == 1;
}

static void stem_fill(StemDev& p, const PPConvDesc& d) {
  p.N = d.N; p.H = d.H; p.W = d.W;
  p.tiles_h = (d.H + kStemTH - 1) / kStemTH;
  p.tiles_w = (d.W + kStemTW - 1) / kStemTW;
  p.num_tiles = d.N * p.tiles_h * p.tiles_w;
}

int stem_grid(const PPConvDesc& d) {
  StemDev p;
  stem_fill(p, d);
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  const int cap = 2 * sms;   // persistent: two CTAs per SM
  return p.num_tiles < cap ? p.num_tiles : cap;
}

// zero-padded input halo of one tile as fp32: s_in[row][col*3 + c], input pixel (p0 - 1 + row, q0 - 1 + col)
__device__ __forceinline__ void stem_load_halo(const __nv_bfloat16* __restrict__ x, float* s_in, int n, int p0, int q0,
                                               int H, int W) {
  const __nv_bfloat16* xi = x + (size_t)n * H * W * 3;
  for (int idx = threadIdx.x; idx < kStemHaloRows * 102; idx += kStemThreads) {
    const int row = idx / 102;
    const int rem = idx - row * 102;
    const int col = rem / 3;
    const int ih = p0 - 1 + row, iw = q0 - 1 + col;
    float v = 0.0f;
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = __bfloat162float(xi[((size_t)ih * W + iw) * 3 + (rem - col * 3)]);
    s_in[row * kStemRowF + rem] = v;
  }
}

__device__ __forceinline__ void stem_tile_coords(const StemDev& p, int tile, int& n, int& p0, int& q0) {
  const int tw = tile % p.tiles_w;
  const int th = (tile / p.tiles_w) % p.tiles_h;
  n = tile / (p.tiles_w * p.tiles_h);
  p0 = th * kStemTH;
  q0 = tw * kStemTW;
}

// 20 consecutive floats (5 broadcast 128-bit loads) of a halo row starting at a 16-byte aligned offset
__device__ __forceinline__ void stem_load_row(const float* s_row, float (&in)[20]) {
  const float4* p4 = reinterpret_cast<const float4*>(s_row);
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const float4 v = p4[i];
    in[4 * i] = v.x; in[4 * i + 1] = v.y; in[4 * i + 2] = v.z; in[4 * i + 3] = v.w;
  }
}

__global__ void __launch_bounds__(kStemThreads, 2)
stem_fprop_kernel(const __grid_constant__ StemDev p, const __nv_bfloat16* __restrict__ x,
                  const __nv_bfloat16* __restrict__ wf /*[64][27]*/) {
  __shared__ __align__(16) float s_in[kStemHaloRows * kStemRowF];
  __shared__ float s_stat[8][4][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = 2 * lane;   // this lane's channels: c0, c0 + 1
  float w0[kStemK], w1[kStemK];
#pragma unroll
  for (int k = 0; k < kStemK; ++k) {
    w0[k] = __bfloat162float(wf[c0 * kStemK + k]);
    w1[k] = __bfloat162float(wf[(c0 + 1) * kStemK + k]);
  }
  const float a0 = p.scale ? p.scale[c0] : 1.0f, a1 = p.scale ? p.scale[c0 + 1] : 1.0f;
  const float b0 = p.shift ? p.shift[c0] : 0.0f, b1 = p.shift ? p.shift[c0 + 1] : 0.0f;
  const bool affine = p.scale != nullptr || p.shift != nullptr;
  float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
  for (int i = threadIdx.x; i < kStemHaloRows * kStemRowF; i += kStemThreads) s_in[i] = 0.0f;   // incl. row padding
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    int n, p0, q0;
    stem_tile_coords(p, tile, n, p0, q0);
    __syncthreads();   // previous tile's readers are done (and the zero fill above)
    stem_load_halo(x, s_in, n, p0, q0, p.H, p.W);
    __syncthreads();
    const int pr = p0 + warp;
    if (pr >= p.H) continue;   // warp-uniform; the barriers above are reached by every warp each iteration
    const size_t row_base = ((size_t)n * p.H + pr) * p.W;
#pragma unroll 1
    for (int s = 0; s < kStemTW / 4; ++s) {
      const int qs = q0 + 4 * s;
      if (qs >= p.W) break;
      float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        float in[20];
        stem_load_row(s_in + (warp + dh) * kStemRowF + 12 * s, in);
#pragma unroll
        for (int px = 0; px < 4; ++px)
#pragma unroll
          for (int dw = 0; dw < 3; ++dw)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float v = in[(px + dw) * 3 + c];
              acc0[px] = fmaf(v, w0[(dh * 3 + dw) * 3 + c], acc0[px]);
              acc1[px] = fmaf(v, w1[(dh * 3 + dw) * 3 + c], acc1[px]);
            }
      }
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        const int q = qs + px;
        if (q < p.W) {
          float z0 = acc0[px], z1 = acc1[px];
          s1a += z0; s1b += z1;
          s2a = fmaf(z0, z0, s2a); s2b = fmaf(z1, z1, s2b);
          if (affine) { z0 = fmaf(z0, a0, b0); z1 = fmaf(z1, a1, b1); }
          if (p.relu) { z0 = fmaxf(z0, 0.0f); z1 = fmaxf(z1, 0.0f); }
          const size_t o = (row_base + q) * kStemO + c0;
          if (p.out_f32) {
            *reinterpret_cast<float2*>(reinterpret_cast<float*>(p.out) + o) = make_float2(z0, z1);
          } else {
            *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o) =
                __floats2bfloat162_rn(z0, z1);
          }
        }
      }
    }
  }
  if (p.stats_partial) {
    s_stat[warp][0][lane] = s1a; s_stat[warp][1][lane] = s1b;
    s_stat[warp][2][lane] = s2a; s_stat[warp][3][lane] = s2b;
    __syncthreads();
    if (threadIdx.x < 128) {
      // thread t: statistic (t / 64), channel (t % 64) = 2*lane' + parity
      const int stat = threadIdx.x >> 6, ch = threadIdx.x & 63;
      const int slot = stat * 2 + (ch & 1), ln = ch >> 1;
      float acc = 0.0f;
#pragma unroll
      for (int w = 0; w < 8; ++w) acc += s_stat[w][slot][ln];
      p.stats_partial[(size_t)blockIdx.x * 2 * kStemO + stat * kStemO + ch] = acc;
    }
  }
}

// partial[blockIdx.x][o][27] = sum over the block's pixels of dz[pixel][o] * patch[pixel][k],  k = (dh*3 + dw)*3 + c
__global__ void __launch_bounds__(kStemThreads, 2)
stem_wgrad_kernel(const __grid_constant__ StemDev p, const __nv_bfloat16* __restrict__ x,
                  const __nv_bfloat16* __restrict__ dz /*[N*H*W, 64]*/, float* __restrict__ partial) {
  __shared__ __align__(16) float s_in[kStemHaloRows * kStemRowF];
  __shared__ __align__(16) __nv_bfloat16 s_dz[kStemTH * kStemTW * kStemO];   // the tile's dz, [row][px][64]: 32 KiB
  __shared__ float s_red[kStemO * kStemK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = 2 * lane;
  float g0[kStemK], g1[kStemK];
#pragma unroll
  for (int k = 0; k < kStemK; ++k) g0[k] = g1[k] = 0.0f;
  for (int i = threadIdx.x; i < kStemHaloRows * kStemRowF; i += kStemThreads) s_in[i] = 0.0f;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    int n, p0, q0;
    stem_tile_coords(p, tile, n, p0, q0);
    __syncthreads();
    stem_load_halo(x, s_in, n, p0, q0, p.H, p.W);
    // dz of the tile through shared memory: eight independent 128-bit loads per thread are in flight at once
    // (a per-strip global load would expose its latency 8 times per row); pixels outside the image read as 0
    for (int idx = threadIdx.x; idx < kStemTH * kStemTW * 8; idx += kStemThreads) {
      const int row = idx >> 8, px = (idx >> 3) & 31, part = idx & 7;
      const int pr_ = p0 + row, q_ = q0 + px;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (pr_ < p.H && q_ < p.W)
        v = *reinterpret_cast<const uint4*>(dz + (((size_t)n * p.H + pr_) * p.W + q_) * kStemO + part * 8);
      reinterpret_cast<uint4*>(s_dz)[idx] = v;
    }
    __syncthreads();
    const int pr = p0 + warp;
    if (pr >= p.H) continue;
#pragma unroll 1
    for (int s = 0; s < kStemTW / 4; ++s) {
      const int qs = q0 + 4 * s;
      if (qs >= p.W) break;
      float d0[4], d1[4];
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        const float2 f = __bfloat1622float2(
            *reinterpret_cast<const __nv_bfloat162*>(s_dz + ((warp * kStemTW) + 4 * s + px) * kStemO + c0));
        d0[px] = f.x; d1[px] = f.y;
      }
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        float in[20];
        stem_load_row(s_in + (warp + dh) * kStemRowF + 12 * s, in);
#pragma unroll
        for (int px = 0; px < 4; ++px)
#pragma unroll
          for (int dw = 0; dw < 3; ++dw)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float v = in[(px + dw) * 3 + c];
              g0[(dh * 3 + dw) * 3 + c] = fmaf(v, d0[px], g0[(dh * 3 + dw) * 3 + c]);
              g1[(dh * 3 + dw) * 3 + c] = fmaf(v, d1[px], g1[(dh * 3 + dw) * 3 + c]);
            }
      }
    }
  }
  // combine the 8 warps in warp order (fixed summation order => deterministic)
  for (int w = 0; w < 8; ++w) {
    __syncthreads();
    if (warp == w) {
#pragma unroll
      for (int k = 0; k < kStemK; ++k) {
        if (w == 0) {
          s_red[c0 * kStemK + k] = g0[k];
          s_red[(c0 + 1) * kStemK + k] = g1[k];
        } else {
          s_red[c0 * kStemK + k] += g0[k];
          s_red[(c0 + 1) * kStemK + k] += g1[k];
        }
      }
    }
  }
  __syncthreads();
  float* dst = partial + (size_t)blockIdx.x * kStemO * kStemK;
  for (int i = threadIdx.x; i < kStemO * kStemK; i += kStemThreads) dst[i] = s_red[i];
}

int stem_fprop(const PPConvDesc& d, const void* x, const void* wf, const TapEpilogue& e, cudaStream_t s) {
  StemDev p;
  stem_fill(p, d);
  p.out = e.out; p.out_f32 = e.out_f32; p.scale = e.scale; p.shift = e.shift; p.relu = e.relu;
  p.stats_partial = e.stats_partial;
  stem_fprop_kernel<<<stem_grid(d), kStemThreads, 0, s>>>(p, (const __nv_bfloat16*)x, (const __nv_bfloat16*)wf);
  PP_POST_LAUNCH();
  return PP_OK;
}

// partial: [stem_grid(d)][64][27] floats
int stem_wgrad(const PPConvDesc& d, const void* x, const void* dz, float* partial, cudaStream_t s) {
  StemDev p;
  stem_fill(p, d);
  p.out = nullptr; p.out_f32 = 0; p.scale = nullptr; p.shift = nullptr; p.relu = 0; p.stats_partial = nullptr;
  stem_wgrad_kernel<<<stem_grid(d), kStemThreads, 0, s>>>(p, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dz,
                                                         partial);
  PP_POST_LAUNCH();
  return PP_OK;
}

}  // namespace pp
