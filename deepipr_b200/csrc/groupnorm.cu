// groupnorm.cu — GroupNorm / InstanceNorm variants of the block's norm:
//   passport blocks  nn.GroupNorm(o // 16, o, affine=False) / nn.InstanceNorm2d(o, affine=False)
//                    models/layers/passportconv2d.py:59-62, passportconv2d_private.py:59-62
//   ConvBlock        nn.GroupNorm(o // 16, o) / nn.InstanceNorm2d(o)            models/layers/conv2d.py:13-16
// used by the scheme-2 attacks (flip_attack.py:215, passport_attack_2.py:144: --norm-type gn).
//
// Statistics are per (sample n, group g) over cpg = O/G channels x HW positions; InstanceNorm is G = O.  The
// normalisation + passport affine + ReLU collapse to  y = relu(a[n,o]*z + b[n,o])  with per-(sample, channel)
// coefficients, and the backward to  dz = k1[n,o]*dy_m + k2[n,o]*z + k3[n,o]  (derivation at gn_bwd_coef_kernel),
// so the passes are the BatchNorm ones (pointwise.cu) with a coefficient row per sample.  All reductions are
// fixed-order: results are run-to-run deterministic.
#include "common.h"
#include "vec8.cuh"

namespace pp {

constexpr int kGnThreads = 256;
constexpr int kGnMaxChunks = 64;

// how many blocks share the HW positions of one sample (>= 2 blocks per SM in total, each with work to do)
int gn_chunks(int N, int HW, int O) {
  const int vec_per_row = O / 8;
  const int row_lanes = kGnThreads / vec_per_row > 0 ? kGnThreads / vec_per_row : 1;
  int want = (2 * 148 + N - 1) / N;
  const int most = (HW + row_lanes - 1) / row_lanes;
  if (want > most) want = most;
  if (want > kGnMaxChunks) want = kGnMaxChunks;
  if (want < 1) want = 1;
  return want;
}

// ---------------------------------------------------------------------------------------------
// per-(sample, channel) sums over the sample's HW positions: partial[(n*chunks + chunk)][2][O]
//   MODE 0: (sum z, sum z^2)                      forward statistics
//   MODE 1: (sum dy_m, sum dy_m*z), dy_m = dy*[a[n,o] z + b[n,o] > 0]   backward reductions
// grid = (chunks, N); thread = (8-channel column, row lane), row lanes combined through shared memory.
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ dy, const void* __restrict__ z, int z_f32, int HW,
                                int O, const float* __restrict__ a, const float* __restrict__ b, int relu,
                                float* __restrict__ partial) {
  extern __shared__ float s_part[];  // [row_lanes][2][O]
  const int vec_per_row = O >> 3;
  const int row_lanes = kGnThreads / vec_per_row > 0 ? kGnThreads / vec_per_row : 1;
  const int col = threadIdx.x % vec_per_row;
  const int rl = threadIdx.x / vec_per_row;
  const int n = blockIdx.y;
  float s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s1[k] = s2[k] = 0.0f;
  if (rl < row_lanes) {
    float ca[8], cb[8];
    if (MODE == 1 && relu) {
      load8_coef(a + (size_t)n * O, col * 8, ca);
      load8_coef(b + (size_t)n * O, col * 8, cb);
    }
    const size_t row0 = (size_t)n * HW;
    for (int r = blockIdx.x * row_lanes + rl; r < HW; r += gridDim.x * row_lanes) {
      const size_t vi = (row0 + r) * vec_per_row + col;
      float zv[8];
      load8(z, z_f32, vi, zv);
      if (MODE == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          s1[k] += zv[k];
          s2[k] = fmaf(zv[k], zv[k], s2[k]);
        }
      } else {
        float g[8];
        load8_bf16(dy, vi, g);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float gm = g[k];
          if (relu && !(fmaf(zv[k], ca[k], cb[k]) > 0.0f)) gm = 0.0f;
          s1[k] += gm;
          s2[k] = fmaf(gm, zv[k], s2[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s_part[(rl * 2 + 0) * O + col * 8 + k] = s1[k];
      s_part[(rl * 2 + 1) * O + col * 8 + k] = s2[k];
    }
  }
  __syncthreads();
  float* dst = partial + ((size_t)n * gridDim.x + blockIdx.x) * 2 * O;
  for (int i = threadIdx.x; i < 2 * O; i += blockDim.x) {
    float acc = 0.0f;
    for (int l = 0; l < row_lanes; ++l) acc += s_part[l * 2 * O + i];
    dst[i] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// forward coefficients, one thread per (n, g):
//   mean = E[z], var = E[z^2] - mean^2 (biased, as F.group_norm / F.instance_norm), invstd = 1/sqrt(var + eps)
//   a[n,o] = gamma[o]*invstd,  b[n,o] = beta[o] - a*mean
// ---------------------------------------------------------------------------------------------
__global__ void gn_fwd_coef_kernel(const float* __restrict__ partial, int chunks, int N, int O, int G, int HW,
                                   float eps, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ save_mean, float* __restrict__ save_invstd,
                                   float* __restrict__ ca, float* __restrict__ cb) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * G) return;
  const int n = idx / G, g = idx - n * G, cpg = O / G;
  double s1 = 0.0, s2 = 0.0;
  for (int k = 0; k < chunks; ++k) {
    const float* p = partial + ((size_t)n * chunks + k) * 2 * O + g * cpg;
    for (int c = 0; c < cpg; ++c) {
      s1 += (double)p[c];
      s2 += (double)p[O + c];
    }
  }
  const double m = (double)cpg * HW;
  const double mean = s1 / m;
  double var = s2 / m - mean * mean;
  if (var < 0.0) var = 0.0;
  const double invstd = 1.0 / sqrt(var + (double)eps);
  save_mean[idx] = (float)mean;
  save_invstd[idx] = (float)invstd;
  for (int c = 0; c < cpg; ++c) {
    const int o = g * cpg + c;
    const float av = (gamma ? gamma[o] : 1.0f) * (float)invstd;
    ca[(size_t)n * O + o] = av;
    cb[(size_t)n * O + o] = (beta ? beta[o] : 0.0f) - av * (float)mean;
  }
}

// backward re-derives the same coefficients from the saved statistics: one thread per (n, o)
__global__ void gn_recoef_kernel(int N, int O, int G, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 const float* __restrict__ mean, const float* __restrict__ invstd,
                                 float* __restrict__ ca, float* __restrict__ cb) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * O) return;
  const int n = (int)(idx / O), o = (int)(idx - (size_t)n * O);
  const int g = o / (O / G);
  const float av = (gamma ? gamma[o] : 1.0f) * invstd[n * G + g];
  ca[idx] = av;
  cb[idx] = (beta ? beta[o] : 0.0f) - av * mean[n * G + g];
}

// ---------------------------------------------------------------------------------------------
// y[r, o] = relu(a[n(r), o]*z[r,o] + b[n(r), o]),  n(r) = r / HW
// ---------------------------------------------------------------------------------------------
__global__ void gn_apply_kernel(const void* __restrict__ z, int z_f32, size_t nvec, int O, int HW,
                                const float* __restrict__ a, const float* __restrict__ b, int relu,
                                __nv_bfloat16* __restrict__ y) {
  const int vec_per_row = O >> 3;
  const size_t vec_per_sample = (size_t)vec_per_row * HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / vec_per_sample;
    const int ch = (int)(i % vec_per_row) << 3;
    float v[8], ca[8], cb[8];
    load8(z, z_f32, i, v);
    load8_coef(a + n * O, ch, ca);
    load8_coef(b + n * O, ch, cb);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      v[k] = fmaf(v[k], ca[k], cb[k]);
      if (relu) v[k] = fmaxf(v[k], 0.0f);
    }
    store8_bf16(y, i, v);
  }
}

// ---------------------------------------------------------------------------------------------
// backward coefficients, one thread per (n, g).  With zhat = (z - mu)*is, dzhat = gamma[o]*dy_m and the group's
// m = cpg*HW elements:   dz = is*(dzhat - mean_S(dzhat) - zhat*mean_S(dzhat*zhat))
//   s1[c] = sum_hw dy_m,  s2[c] = sum_hw dy_m*z   (gn_stats_kernel<1>)
//   A1 = sum_c gamma[c]*s1[c] / m          A2 = sum_c gamma[c]*is*(s2[c] - mu*s1[c]) / m
//   k1[n,o] = is*gamma[o]    k2[n,o] = -is^2*A2    k3[n,o] = -is*A1 + is^2*A2*mu
// and the per-sample terms of the affine gradients: contrib[n][0][o] = is*(s2 - mu*s1) (dgamma),
// contrib[n][1][o] = s1 (dbeta), summed over n by gn_dparam_kernel.
// ---------------------------------------------------------------------------------------------
__global__ void gn_bwd_coef_kernel(const float* __restrict__ partial, int chunks, int N, int O, int G, int HW,
                                   const float* __restrict__ gamma, const float* __restrict__ mean,
                                   const float* __restrict__ invstd, float* __restrict__ contrib,
                                   float* __restrict__ k1, float* __restrict__ k2, float* __restrict__ k3) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * G) return;
  const int n = idx / G, g = idx - n * G, cpg = O / G;
  const double mu = mean[idx], is = invstd[idx];
  double A1 = 0.0, A2 = 0.0;
  for (int c = 0; c < cpg; ++c) {
    const int o = g * cpg + c;
    double s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < chunks; ++k) {
      const float* p = partial + ((size_t)n * chunks + k) * 2 * O;
      s1 += (double)p[o];
      s2 += (double)p[O + o];
    }
    const double gam = gamma ? (double)gamma[o] : 1.0;
    const double t = is * (s2 - mu * s1);
    A1 += gam * s1;
    A2 += gam * t;
    contrib[((size_t)n * 2 + 0) * O + o] = (float)t;
    contrib[((size_t)n * 2 + 1) * O + o] = (float)s1;
  }
  const double m = (double)cpg * HW;
  A1 /= m;
  A2 /= m;
  for (int c = 0; c < cpg; ++c) {
    const int o = g * cpg + c;
    const double gam = gamma ? (double)gamma[o] : 1.0;
    k1[(size_t)n * O + o] = (float)(is * gam);
    k2[(size_t)n * O + o] = (float)(-is * is * A2);
    k3[(size_t)n * O + o] = (float)(-is * A1 + is * is * A2 * mu);
  }
}

// dgamma[o] = sum_n contrib[n][0][o], dbeta[o] = sum_n contrib[n][1][o]; block (32 channels, 32 sample slices)
__global__ void __launch_bounds__(1024) gn_dparam_kernel(const float* __restrict__ contrib, int N, int O,
                                                         float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                         int flags) {
  __shared__ double sh1[32][33], sh2[32][33];
  const int o = blockIdx.x * 32 + threadIdx.x;
  const int slice = threadIdx.y;
  double a1 = 0.0, a2 = 0.0;
  if (o < O) {
    for (int n = slice; n < N; n += 32) {
      a1 += (double)contrib[((size_t)n * 2 + 0) * O + o];
      a2 += (double)contrib[((size_t)n * 2 + 1) * O + o];
    }
  }
  sh1[slice][threadIdx.x] = a1;
  sh2[slice][threadIdx.x] = a2;
  __syncthreads();
  if (slice != 0 || o >= O) return;
  double s1 = 0.0, s2 = 0.0;
  for (int k = 0; k < 32; ++k) {
    s1 += sh1[k][threadIdx.x];
    s2 += sh2[k][threadIdx.x];
  }
  dgamma[o] = (flags & PP_FLAG_ACC_DGAMMA) ? dgamma[o] + (float)s1 : (float)s1;
  dbeta[o] = (flags & PP_FLAG_ACC_DBETA) ? dbeta[o] + (float)s2 : (float)s2;
}

// dz[r,o] = k1[n,o]*dy_m + k2[n,o]*z + k3[n,o]
__global__ void gn_dz_kernel(const __nv_bfloat16* __restrict__ dy, const void* __restrict__ z, int z_f32, size_t nvec,
                             int O, int HW, const float* __restrict__ a, const float* __restrict__ b, int relu,
                             const float* __restrict__ k1, const float* __restrict__ k2, const float* __restrict__ k3,
                             __nv_bfloat16* __restrict__ dz) {
  const int vec_per_row = O >> 3;
  const size_t vec_per_sample = (size_t)vec_per_row * HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / vec_per_sample;
    const int ch = (int)(i % vec_per_row) << 3;
    float zv[8], g[8], ca[8], cb[8], c1[8], c2[8], c3[8], out[8];
    load8(z, z_f32, i, zv);
    load8_bf16(dy, i, g);
    load8_coef(a + n * O, ch, ca);
    load8_coef(b + n * O, ch, cb);
    load8_coef(k1 + n * O, ch, c1);
    load8_coef(k2 + n * O, ch, c2);
    load8_coef(k3 + n * O, ch, c3);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float gm = g[k];
      if (relu && !(fmaf(zv[k], ca[k], cb[k]) > 0.0f)) gm = 0.0f;
      out[k] = fmaf(c1[k], gm, fmaf(c2[k], zv[k], c3[k]));
    }
    store8_bf16(dz, i, out);
  }
}

static inline int gn_grid_for(size_t work_items, int threads, int max_blocks) {
  size_t b = (work_items + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > (size_t)max_blocks) b = max_blocks;
  return (int)b;
}

static int gn_check(const PPConvDesc& d) {
  PP_REQUIRE(d.groups > 0 && d.O % d.groups == 0, PP_EBADSHAPE, "group norm: O=%d is not divisible by groups=%d", d.O,
             d.groups);
  PP_REQUIRE(d.O % 8 == 0 && d.O / 8 <= kGnThreads, PP_EBADSHAPE, "group norm needs O%%8==0 and O<=2048 (O=%d)", d.O);
  return PP_OK;
}

// forward: statistics of z -> save_mean/save_invstd [N*G], coefficients [N*O], y
int launch_gn_fwd(const PPConvDesc& d, int HW, const void* z, const float* gamma, const float* beta, float* save_mean,
                  float* save_invstd, float* ca, float* cb, float* partial, __nv_bfloat16* y, cudaStream_t s) {
  PP_TRY(gn_check(d));
  const int chunks = gn_chunks(d.N, HW, d.O);
  const int vec_per_row = d.O / 8;
  const int row_lanes = kGnThreads / vec_per_row;
  const size_t smem = (size_t)row_lanes * 2 * d.O * sizeof(float);
  gn_stats_kernel<0><<<dim3(chunks, d.N), kGnThreads, smem, s>>>(nullptr, z, d.z_f32, HW, d.O, nullptr, nullptr, 0,
                                                                  partial);
  PP_POST_LAUNCH();
  const int ng = d.N * d.groups;
  gn_fwd_coef_kernel<<<(ng + 127) / 128, 128, 0, s>>>(partial, chunks, d.N, d.O, d.groups, HW, d.eps, gamma, beta,
                                                      save_mean, save_invstd, ca, cb);
  PP_POST_LAUNCH();
  const size_t nvec = (size_t)d.N * HW * vec_per_row;
  gn_apply_kernel<<<gn_grid_for(nvec, 256, 148 * 8), 256, 0, s>>>(z, d.z_f32, nvec, d.O, HW, ca, cb, d.relu, y);
  PP_POST_LAUNCH();
  return PP_OK;
}

// backward, part 1: coefficients + reductions + dgamma/dbeta
int launch_gn_bwd_reduce(const PPConvDesc& d, int HW, const __nv_bfloat16* dy, const void* z, const float* gamma,
                         const float* beta, const float* save_mean, const float* save_invstd, float* ca, float* cb,
                         float* k1, float* k2, float* k3, float* partial, float* contrib, float* dgamma, float* dbeta,
                         cudaStream_t s) {
  PP_TRY(gn_check(d));
  const int chunks = gn_chunks(d.N, HW, d.O);
  const int vec_per_row = d.O / 8;
  const int row_lanes = kGnThreads / vec_per_row;
  const size_t smem = (size_t)row_lanes * 2 * d.O * sizeof(float);
  const size_t no = (size_t)d.N * d.O;
  gn_recoef_kernel<<<(unsigned)((no + 255) / 256), 256, 0, s>>>(d.N, d.O, d.groups, gamma, beta, save_mean,
                                                                 save_invstd, ca, cb);
  PP_POST_LAUNCH();
  gn_stats_kernel<1><<<dim3(chunks, d.N), kGnThreads, smem, s>>>(dy, z, d.z_f32, HW, d.O, ca, cb, d.relu, partial);
  PP_POST_LAUNCH();
  const int ng = d.N * d.groups;
  gn_bwd_coef_kernel<<<(ng + 127) / 128, 128, 0, s>>>(partial, chunks, d.N, d.O, d.groups, HW, gamma, save_mean,
                                                      save_invstd, contrib, k1, k2, k3);
  PP_POST_LAUNCH();
  gn_dparam_kernel<<<(d.O + 31) / 32, dim3(32, 32), 0, s>>>(contrib, d.N, d.O, dgamma, dbeta, d.flags);
  PP_POST_LAUNCH();
  return PP_OK;
}

// backward, part 2: dz
int launch_gn_dz(const PPConvDesc& d, int HW, const __nv_bfloat16* dy, const void* z, const float* ca,
                 const float* cb, const float* k1, const float* k2, const float* k3, __nv_bfloat16* dz,
                 cudaStream_t s) {
  const size_t nvec = (size_t)d.N * HW * (d.O / 8);
  gn_dz_kernel<<<gn_grid_for(nvec, 256, 148 * 8), 256, 0, s>>>(dy, z, d.z_f32, nvec, d.O, HW, ca, cb, d.relu, k1, k2,
                                                               k3, dz);
  PP_POST_LAUNCH();
  return PP_OK;
}

}  // namespace pp
