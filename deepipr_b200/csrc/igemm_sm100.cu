// igemm_sm100.cu — the tensor-core kernels of the passport-conv path (sm_100a only).
//
//   tapgemm_kernel : persistent, warp-specialised implicit-GEMM convolution
//                    A (activation patches)  : TMA im2col -> 128B-swizzled smem, K-major
//                    B (weights)             : TMA tiled  -> 128B-swizzled smem, K-major
//                    D                       : tcgen05.mma kind::f16 (bf16 x bf16 -> fp32) into TMEM,
//                                              two accumulator buffers so the epilogue of tile i overlaps
//                                              the main loop of tile i+1
//                    epilogue                : tcgen05.ld -> registers -> per-channel sum / sum-of-squares
//                                              (warp-shuffle transpose-reduce) and/or affine+ReLU -> 128-bit stores
//                    serves conv fprop (passportconv2d.py:218) and its data gradient.
//   wgrad_kernel   : weight gradient computed transposed (rows = (tap, channel), columns = output channel),
//                    both operands MN-major (pixels are the reduction dimension), split over pixel ranges;
//                    fp32 partial tiles go to a workspace reduced by pointwise.cu:wgrad_finalize
//                    (deterministic order, no float atomics).
//
// ptx.cuh is included by this translation unit only (it defines a __device__ variable).
#include "common.h"
#include "ptx.cuh"

#include <stdlib.h>

#include <mutex>

namespace pp {

// ------------------------------------------------------------------------------------------------
// driver entry points for tensor-map encoding (resolved at run time: the library must load on a
// machine without libcuda.so.1, e.g. the CPU-only CI container)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn g_encode_tiled = nullptr;
static EncodeIm2colFn g_encode_im2col = nullptr;
static int g_driver_version = 0;
static std::once_flag g_encode_once;

static int resolve_encoders() {
  std::call_once(g_encode_once, [] {
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
    fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
    cudaDriverGetVersion(&g_driver_version);
  });
  if (!g_encode_tiled || !g_encode_im2col) {
    set_error("cuTensorMapEncode{Tiled,Im2col} not available from the driver");
    return PP_ENODEVICE;
  }
  return PP_OK;
}

// 2-D row-major matrix [rows, cols] of bf16 (esz 2) or fp32 (esz 4); box = box_rows x 128 bytes, 128B swizzle.
// `mn32`: SWIZZLE_128B_ATOM_32B instead of SWIZZLE_128B — the only shared-memory layout tcgen05 accepts for MN-major
// operands of 4-byte types (UMMA layout type 128B_BASE32B: 32-byte chunks XORed with the row index mod 4).
static int make_map_2d(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, int esz = 2,
                       bool mn32 = false) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * (uint64_t)esz};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(m, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                              const_cast<void*>(ptr), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              mn32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu box_rows=%u", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, box_rows);
    return PP_ELAUNCH;
  }
  return PP_OK;
}

// NHWC activation (bf16, or fp32 when g.tf32) in im2col mode: 128 bytes of channels x `pixels` traversal positions.
static int make_map_im2col(CUtensorMap* m, const void* ptr, const TapGemm& g, uint32_t pixels, bool mn32 = false) {
  const uint64_t esz = g.tf32 ? 4 : 2;
  cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.N};
  cuuint64_t strides[3] = {(cuuint64_t)g.C * esz, (cuuint64_t)g.W * g.C * esz, (cuuint64_t)g.H * g.W * g.C * esz};
  int lower[2] = {g.base_w, g.base_h};
  int upper[2] = {g.upper_w, g.upper_h};
  cuuint32_t estr[4] = {1, (cuuint32_t)g.step_w, (cuuint32_t)g.step_h, 1};
  CUresult r = g_encode_im2col(m, g.tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                               const_cast<void*>(ptr), dims, strides, lower, upper, (cuuint32_t)(128 / esz), pixels, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE,
                               mn32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeIm2col failed (%d) NHWC=%d,%d,%d,%d lower=%d,%d upper=%d,%d step=%d,%d", (int)r,
              g.N, g.H, g.W, g.C, g.base_w, g.base_h, g.upper_w, g.upper_h, g.step_w, g.step_h);
    return PP_ELAUNCH;
  }
  // Same driver workaround CUTLASS applies (cute/atom/copy_traits_sm90_im2col.hpp): on drivers <= 13.1 an
  // im2col descriptor of a tensor smaller than 128 KiB must have bit 21 of its second 64-bit word cleared.
  if (g_driver_version <= 13010) {
    uint64_t bytes = (uint64_t)g.N * g.H * g.W * g.C * esz;
    if (bytes < 131072) reinterpret_cast<uint64_t*>(m)[1] &= ~(1ull << 21);
  }
  return PP_OK;
}

// When a tile of `pixels` traversal positions is a whole number of image rows / images (stride 1, power-of-two
// feature maps: every CIFAR-shaped layer), the same shared-memory image is produced by an ordinary TILED 4-D box
// {64 ch, Q, rows, images} whose start coordinate is shifted by the filter tap; out-of-bound parts (padding,
// batch tail) are zero-filled exactly as in im2col mode.  Returns false when the geometry does not fit.
static bool tiled_box_for(const TapGemm& g, int pixels, int* bw, int* bh, int* bn) {
  if (g.step_h != 1 || g.step_w != 1) return false;
  const int PQ = g.P * g.Q;
  if (g.Q > pixels || pixels % g.Q != 0) return false;
  if (!((PQ % pixels == 0) || (pixels % PQ == 0))) return false;
  *bw = g.Q;
  if (PQ >= pixels) { *bh = pixels / g.Q; *bn = 1; }
  else { *bh = g.P; *bn = pixels / PQ; }
  if (*bw > 256 || *bh > 256 || *bn > 256) return false;
  return true;
}

static int make_map_tiled4d(CUtensorMap* m, const void* ptr, const TapGemm& g, int bw, int bh, int bn) {
  const uint64_t esz = g.tf32 ? 4 : 2;
  cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.N};
  cuuint64_t strides[3] = {(cuuint64_t)g.C * esz, (cuuint64_t)g.W * g.C * esz, (cuuint64_t)g.H * g.W * g.C * esz};
  cuuint32_t box[4] = {(cuuint32_t)(128 / esz), (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode_tiled(m, g.tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                              const_cast<void*>(ptr), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4d) failed (%d) NHWC=%d,%d,%d,%d box=%d,%d,%d", (int)r, g.N, g.H, g.W, g.C, bw,
              bh, bn);
    return PP_ELAUNCH;
  }
  return PP_OK;
}

static bool g_prefer_tiled = true;   // PP_TMA_IM2COL_ONLY=1 forces the im2col-mode loads (A/B comparison)
static bool prefer_tiled() {
  static bool init = false;
  if (!init) {
    const char* e = getenv("PP_TMA_IM2COL_ONLY");
    g_prefer_tiled = !(e && e[0] == '1');
    const char* t = getenv("PP_TMA_TILED");   // default: im2col mode (validated on every geometry); =1 tries tiled boxes
    g_prefer_tiled = (t && t[0] == '1');
    init = true;
  }
  return g_prefer_tiled;
}

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
template <int OFF>
__device__ __forceinline__ void bfly_step(float (&v)[32], int lane) {
  const bool upper = (lane & OFF) != 0;
#pragma unroll
  for (int i = 0; i < OFF; ++i) {
    const float send = upper ? v[i] : v[i + OFF];
    const float keep = upper ? v[i + OFF] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
  }
}
// Each lane holds one row of 32 columns; afterwards lane L holds the sum of column L over the 32 rows.
__device__ __forceinline__ float warp_column_sum(float (&v)[32], int lane) {
  bfly_step<16>(v, lane);
  bfly_step<8>(v, lane);
  bfly_step<4>(v, lane);
  bfly_step<2>(v, lane);
  bfly_step<1>(v, lane);
  return v[0];
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// 8 bf16 + 8 bf16 -> 8 bf16 (fp32 add, round to nearest)
__device__ __forceinline__ uint4 add_bf16x8(uint4 a, uint4 b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 fa = __bfloat1622float2(pa[i]), fb = __bfloat1622float2(pb[i]);
    pr[i] = __floats2bfloat162_rn(fa.x + fb.x, fa.y + fb.y);
  }
  return r;
}

struct TapGemmDev {
  int M;            // N * P * Q
  int P, Q, PQ;
  int base_h, base_w, step_h, step_w;
  int C;            // channels of the activation (K per tap)
  int ntaps;
  int Nout;
  int num_m_tiles, num_n_tiles;
  int out_H, out_W, out_sh, out_sw, out_ph, out_pw, out_identity;
  int a_tiled;      // 1: activation tiles are loaded with a tiled 4-D box instead of im2col mode
  int dbg;          // bring-up switches (PP_DEBUG env): 1 skip global stores, 2 skip statistics, 4 skip TMEM loads,
                    // 8 skip the MMA instructions, 16 skip the TMA loads (results are garbage; timing experiments only)
  int8_t tap_dh[kMaxTaps], tap_dw[kMaxTaps];
  int tap_kofs[kMaxTaps];
  // epilogue
  void* out;
  int out_f32;
  const float* scale;
  const float* shift;
  int relu;
  float* stats_partial;
  const __nv_bfloat16* add_src;   // see TapEpilogue::add_src
};

constexpr int kBM = 128;       // output pixels per tile (UMMA M)
constexpr int kBK = 64;        // channels per pipeline stage (one 128-byte swizzle row of bf16)
constexpr int kThreads = 192;  // warp 0 TMA, warp 1 MMA + TMEM owner, warps 2..5 epilogue
constexpr int kMaxStatsN = 512;   // widest output for which the fused column statistics are kept in smem

constexpr int kResMaxSteps = 9;   // resident-weight variant: at most 9 (tap, channel-chunk) blocks of BN x 64

// RESB: the whole weight matrix (one n tile, <= 9 k-steps) is loaded into shared memory once per CTA and stays
// there for all of its tiles; the pipeline stages then carry activation tiles only.  For the 64-channel layers
// this removes a third of the L2->smem traffic and the 148-way hot spot on the same 72 KiB of weights.
template <int BN, bool RESB = false>
struct FwdCfg {
  static constexpr int kStageA = kBM * kBK * 2;  // 16 KiB
  static constexpr int kStageB = BN * kBK * 2;
  static constexpr int kStageBytes = RESB ? kStageA : kStageA + kStageB;
  static constexpr int kResBytes = RESB ? kResMaxSteps * kStageB : 0;
  static constexpr int kStages = RESB ? 8 : ((BN == 64) ? 8 : (BN == 128 ? 6 : 4));
  static constexpr int kTmemCols = BN == 192 ? 512 : 2 * BN;  // two accumulator buffers (allocations are powers of 2)
  static constexpr int kScratch = 2 * BN * 4 /*a,b*/ + 4 * 2 * BN * 4 /*per-warp column sums*/ +
                                  2 * kMaxStatsN * 4 /*per-CTA running column sums*/ +
                                  4 * 4096 /*per-warp store staging tiles*/;
  static constexpr int kSmemBytes =
      1024 /*align slack*/ + kResBytes + kStages * kStageBytes + kScratch + 256 /*barriers*/;
  static_assert(kSmemBytes <= 232448, "tapgemm stage configuration exceeds the 227 KiB of shared memory per CTA");
};

// TF32: activations and weights are fp32 in memory and enter the tensor core as TF32 (tcgen05.mma kind::tf32, K = 8
// per instruction).  A pipeline stage is still kBM (or BN) rows of ONE 128-byte swizzle row — 32 fp32 channels instead
// of 64 bf16 ones — so shared-memory layout, descriptors and the 32-byte K advance are unchanged; only the channel
// count per stage (kBKe), the instruction kind and its descriptor differ.
template <int BN, bool RESB, bool TF32 = false>
__global__ void __launch_bounds__(kThreads, 1)
tapgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ TapGemmDev p) {
  using Cfg = FwdCfg<BN, RESB>;
  constexpr int kBKe = TF32 ? 32 : kBK;        // channels (elements) per pipeline stage
  constexpr int kAccStride = BN == 192 ? 256 : BN;   // TMEM column distance of the two accumulator buffers
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* res_base = smem;                       // [ksteps][BN x 64] resident weight blocks (RESB only)
  uint8_t* stage_base = smem + Cfg::kResBytes;
  float* s_a = reinterpret_cast<float*>(stage_base + Cfg::kStages * Cfg::kStageBytes);
  float* s_b = s_a + BN;
  float* s_red = s_b + BN;            // [4 warps][2][BN]
  float* s_acc = s_red + 4 * 2 * BN;  // [2][Nout] running per-CTA column sums (sum z, sum z^2)
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_acc + 2 * kMaxStatsN);  // [4 warps][4 KiB], 16-byte aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + 4 * 4096);
  uint64_t* full = bars;                       // [kStages]
  uint64_t* empty = bars + Cfg::kStages;       // [kStages]
  uint64_t* tfull = bars + 2 * Cfg::kStages;   // [2]
  uint64_t* tempty = tfull + 2;                // [2]
  uint64_t* bres = tempty + 2;                 // [1] resident weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int kchunks = p.C / kBKe;
  const int ksteps = p.ntaps * kchunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);  // one arrival per epilogue warp
    }
    mbar_init(bres, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The whole warp runs the loop (warp-uniform control flow keeps addresses / descriptors in uniform registers);
    // one elected lane issues the bulk-tensor copies.
    if (RESB) {
      if (elect_one()) {
        // all weight blocks once: block ks = (tap t, channel chunk kc) in main-loop order
        mbar_arrive_expect_tx(bres, (uint32_t)ksteps * Cfg::kStageB);
        for (int t = 0; t < p.ntaps; ++t)
          for (int kc = 0; kc < kchunks; ++kc)
            tma_load_2d(&tmB, bres, res_base + (t * kchunks + kc) * Cfg::kStageB, p.tap_kofs[t] + kc * kBKe, 0);
      }
      __syncwarp();
    }
    int stage = 0;
    uint32_t phase = 0;
    const int dbg = p.dbg;
    const int a_tiled = p.a_tiled;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.num_n_tiles;
      const int m_tile = tile / p.num_n_tiles;
      const int m0 = m_tile * kBM;
      const int img = m0 / p.PQ;
      const int rem = m0 - img * p.PQ;
      const int p0 = rem / p.Q;
      const int q0 = rem - p0 * p.Q;
      const int cw = p.base_w + q0 * p.step_w;
      const int ch = p.base_h + p0 * p.step_h;
      for (int t = 0; t < p.ntaps; ++t) {
        const int dw = p.tap_dw[t];
        const int dh = p.tap_dh[t];
        const int kofs = p.tap_kofs[t];
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1, 100 + stage);
          if (elect_one()) {
            uint8_t* sa = stage_base + stage * Cfg::kStageBytes;
            uint8_t* sb = sa + Cfg::kStageA;
            if (dbg & 16) {   // bring-up: no loads at all, just hand the (stale) stage to the MMA warp
              mbar_arrive(&full[stage]);
            } else {
              mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
              if (a_tiled) tma_load_4d(&tmA, &full[stage], sa, kc * kBKe, cw + dw, ch + dh, img);
              else tma_load_im2col_4d(&tmA, &full[stage], sa, kc * kBKe, cw, ch, img, (uint16_t)dw, (uint16_t)dh);
              if (!RESB) tma_load_2d(&tmB, &full[stage], sb, kofs + kc * kBKe, n_tile * BN);
            }
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Warp-uniform loop; one elected lane issues tcgen05.mma / tcgen05.commit (descriptors stay in uniform registers,
    // so the instructions issue back to back instead of through per-instruction election loops).
    constexpr uint32_t idesc = TF32 ? make_idesc_tf32(kBM, BN, 0, 0) : make_idesc_bf16(kBM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int dbg = p.dbg;
    const uint32_t stage0 = smem_u32(stage_base);
    const uint32_t res0 = smem_u32(res_base);
    if (RESB) mbar_wait(bres, 0, 250);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1, 200 + acc);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kAccStride;
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full[stage], phase, 300 + stage);
        tc_fence_after();
        const uint32_t sa = stage0 + stage * Cfg::kStageBytes;
        const uint32_t sb = RESB ? res0 + ks * Cfg::kStageB : sa + Cfg::kStageA;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t da = make_smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(sb + k * 32, 16, 1024);
            if (!(dbg & 8)) {
              if (TF32) tc_mma_tf32(d_tmem, da, db, idesc, (ks | k) != 0 ? 1u : 0u);
              else tc_mma_bf16(d_tmem, da, db, idesc, (ks | k) != 0 ? 1u : 0u);
            }
          }
          tc_commit(&empty[stage]);  // frees the smem slot once these MMAs have read it
          if (ks == ksteps - 1) tc_commit(&tfull[acc]);  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue (4 warps, 128 threads = 128 accumulator rows) =====================
    const int ew = warp - 2;          // 0..3: slot in the reduction scratch
    const int quarter = warp & 3;     // TMEM lane quarter this warp may read
    const int row = quarter * 32 + lane;
    const int tid_e = ew * 32 + lane;
    const bool has_affine = (p.scale != nullptr) || (p.shift != nullptr);
    const bool has_stats = p.stats_partial != nullptr;
    int acc = 0;
    uint32_t acc_phase = 0;
    if (has_stats) {
      for (int i = tid_e; i < 2 * p.Nout; i += 128) s_acc[i] = 0.0f;
    }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.num_n_tiles;
      const int m_tile = tile / p.num_n_tiles;
      const int n0 = n_tile * BN;
      const int m = m_tile * kBM + row;
      const bool valid = m < p.M;
      if (has_affine || has_stats) {
        named_bar_sync(1, 128);  // previous tile's readers of s_a / s_red are done
        if (has_affine) {
          for (int i = tid_e; i < BN; i += 128) {
            s_a[i] = p.scale ? __ldg(p.scale + n0 + i) : 1.0f;
            s_b[i] = p.shift ? __ldg(p.shift + n0 + i) : 0.0f;
          }
          named_bar_sync(1, 128);
        }
      }
      size_t out_row = 0;
      if (valid) {
        if (p.out_identity) {
          out_row = (size_t)m;
        } else {
          const int img = m / p.PQ;
          const int rem = m - img * p.PQ;
          const int pp_ = rem / p.Q;
          const int qq_ = rem - pp_ * p.Q;
          out_row = ((size_t)img * p.out_H + (size_t)(pp_ * p.out_sh + p.out_ph)) * p.out_W +
                    (size_t)(qq_ * p.out_sw + p.out_pw);
        }
      }
      const uint32_t valid_mask = __ballot_sync(0xffffffffu, valid);
      mbar_wait(&tfull[acc], acc_phase, 400 + acc);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + acc * kAccStride;
#pragma unroll 1
      for (int j = 0; j < BN / 32; ++j) {
        uint32_t raw[32];
        if (p.dbg & 4) {
#pragma unroll
          for (int i = 0; i < 32; ++i) raw[i] = 0x3f800000u;
        } else {
          tmem_ld_32x32(taddr + j * 32, raw);
          tmem_ld_wait();
        }
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = valid ? __uint_as_float(raw[i]) : 0.0f;
        if (has_stats && !(p.dbg & 2)) {
          float t1[32], t2[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            t1[i] = v[i];
            t2[i] = v[i] * v[i];
          }
          const float s1 = warp_column_sum(t1, lane);
          const float s2 = warp_column_sum(t2, lane);
          s_red[(ew * 2 + 0) * BN + j * 32 + lane] = s1;
          s_red[(ew * 2 + 1) * BN + j * 32 + lane] = s2;
        }
        if (has_affine) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], s_a[j * 32 + i], s_b[j * 32 + i]);
        }
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
        }
        // Store through a per-warp staging tile so that every store instruction writes whole 128-byte lines
        // (a lane owns a ROW of the accumulator; writing it directly would touch 32 different lines per store).
        if (p.dbg & 1) continue;
        if (p.out_f32) {
          float4* st = reinterpret_cast<float4*>(s_stage + ew * 4096);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            st[lane * 8 + (i ^ (lane & 7))] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          __syncwarp();
          float* obase = reinterpret_cast<float*>(p.out) + n0 + j * 32 + (lane & 7) * 4;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + (lane >> 3);
            const float4 val = st[r * 8 + ((lane & 7) ^ (r & 7))];
            const long long orow = __shfl_sync(0xffffffffu, (long long)out_row, r);
            if ((valid_mask >> r) & 1u) *reinterpret_cast<float4*>(obase + (size_t)orow * p.Nout) = val;
          }
          __syncwarp();
        } else {
          uint4* st = reinterpret_cast<uint4*>(s_stage + ew * 4096);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 u;
            u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
            u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
            u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
            u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
            st[lane * 4 + (i ^ ((lane >> 1) & 3))] = u;
          }
          __syncwarp();
          __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(p.out) + n0 + j * 32 + (lane & 3) * 8;
          const __nv_bfloat16* abase = p.add_src ? p.add_src + n0 + j * 32 + (lane & 3) * 8 : nullptr;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int r = it * 8 + (lane >> 2);
            uint4 val = st[r * 4 + ((lane & 3) ^ ((r >> 1) & 3))];
            const long long orow = __shfl_sync(0xffffffffu, (long long)out_row, r);
            if ((valid_mask >> r) & 1u) {
              if (abase) val = add_bf16x8(val, __ldg(reinterpret_cast<const uint4*>(abase + (size_t)orow * p.Nout)));
              *reinterpret_cast<uint4*>(obase + (size_t)orow * p.Nout) = val;
            }
          }
          __syncwarp();
        }
      }
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (has_stats) {
        named_bar_sync(1, 128);
        for (int i = tid_e; i < BN; i += 128) {
          const float a1 = s_red[(0 * 2 + 0) * BN + i] + s_red[(1 * 2 + 0) * BN + i] + s_red[(2 * 2 + 0) * BN + i] +
                           s_red[(3 * 2 + 0) * BN + i];
          const float a2 = s_red[(0 * 2 + 1) * BN + i] + s_red[(1 * 2 + 1) * BN + i] + s_red[(2 * 2 + 1) * BN + i] +
                           s_red[(3 * 2 + 1) * BN + i];
          // column n0+i is only ever touched by this thread (i == column % BN), so no race across tiles
          s_acc[n0 + i] += a1;
          s_acc[p.Nout + n0 + i] += a2;
        }
      }
    }
    if (has_stats) {
      named_bar_sync(1, 128);
      float* dst = p.stats_partial + (size_t)blockIdx.x * 2 * p.Nout;
      for (int i = tid_e; i < 2 * p.Nout; i += 128) dst[i] = s_acc[i];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

bool tapgemm_tcgen05_supported(const TapGemm& g) {
  if (g.C % (g.tf32 ? 32 : 64) != 0 || g.Nout % 64 != 0) return false;
  if (g.ntaps < 1 || g.ntaps > kMaxTaps) return false;
  if (g.base_h < -128 || g.base_h > 127 || g.base_w < -128 || g.base_w > 127) return false;
  if (g.upper_h < -128 || g.upper_h > 127 || g.upper_w < -128 || g.upper_w > 127) return false;
  if (g.step_h < 1 || g.step_h > 8 || g.step_w < 1 || g.step_w > 8) return false;
  for (int t = 0; t < g.ntaps; ++t)
    if (g.tap_dh[t] < 0 || g.tap_dw[t] < 0) return false;
  if ((long long)g.N * g.P * g.Q > 0x7fffffffLL) return false;
  return true;
}

// Output-channel tile: 256 halves the smem bytes read per MAC (an SS-mode 128x128 MMA already needs the full
// 128 B/clk of shared-memory bandwidth); it is used whenever it still yields at least one tile per SM.
static int pick_bn(const TapGemm& g) {
  const long long M = (long long)g.N * g.P * g.Q;
  const long long m_tiles = (M + kBM - 1) / kBM;
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  if (g.Nout % 256 == 0 && m_tiles * (g.Nout / 256) >= sms) return 256;
  // 192 / 384 output channels (AlexNet): 192-wide tiles instead of 3 x 64 / 3 x 128
  if (g.Nout % 192 == 0 && g.Nout % 256 != 0 && m_tiles * (g.Nout / 192) >= sms) return 192;
  if (g.Nout % 128 == 0) return 128;
  return 64;
}

int tapgemm_tcgen05_max_stats_width() { return kMaxStatsN; }

template <int BN, bool RESB, bool TF32 = false>
static int launch_tapgemm(const TapGemm& g, const void* act, const void* B, const TapEpilogue& e, cudaStream_t s) {
  using Cfg = FwdCfg<BN, RESB>;
  CUtensorMap tmA, tmB;
  int bw = 0, bh = 0, bn = 0;
  const bool a_tiled = prefer_tiled() && tiled_box_for(g, kBM, &bw, &bh, &bn);
  if (a_tiled) PP_TRY(make_map_tiled4d(&tmA, act, g, bw, bh, bn));
  else PP_TRY(make_map_im2col(&tmA, act, g, kBM));
  PP_TRY(make_map_2d(&tmB, B, (uint64_t)g.Nout, (uint64_t)g.Ktot, BN, TF32 ? 4 : 2));
  TapGemmDev p;
  p.a_tiled = a_tiled ? 1 : 0;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("PP_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.dbg = dbg;
  }
  p.M = g.N * g.P * g.Q;
  p.P = g.P; p.Q = g.Q; p.PQ = g.P * g.Q;
  p.base_h = g.base_h; p.base_w = g.base_w; p.step_h = g.step_h; p.step_w = g.step_w;
  p.C = g.C; p.ntaps = g.ntaps; p.Nout = g.Nout;
  p.num_m_tiles = (p.M + kBM - 1) / kBM;
  p.num_n_tiles = g.Nout / BN;
  p.out_H = g.out_H; p.out_W = g.out_W; p.out_sh = g.out_sh; p.out_sw = g.out_sw;
  p.out_ph = g.out_ph; p.out_pw = g.out_pw; p.out_identity = g.out_identity;
  for (int t = 0; t < g.ntaps; ++t) {
    p.tap_dh[t] = g.tap_dh[t]; p.tap_dw[t] = g.tap_dw[t]; p.tap_kofs[t] = g.tap_kofs[t];
  }
  p.out = e.out; p.out_f32 = e.out_f32; p.scale = e.scale; p.shift = e.shift; p.relu = e.relu;
  p.stats_partial = e.stats_partial;
  p.add_src = (const __nv_bfloat16*)e.add_src;
  PP_REQUIRE(!e.add_src || !e.out_f32, PP_EUNSUPPORTED, "add_src needs a bf16 output");

  PP_SET_MAX_SMEM_ONCE((tapgemm_kernel<BN, RESB, TF32>), Cfg::kSmemBytes);
  PP_REQUIRE(e.stats_partial == nullptr || g.Nout <= kMaxStatsN, PP_EUNSUPPORTED,
             "fused column statistics support Nout <= %d (Nout=%d)", kMaxStatsN, g.Nout);
  const int grid = tapgemm_tcgen05_grid(g);
  prof_begin(PROF_TAPGEMM, 2.0 * (double)p.M * g.Nout * g.ntaps * g.C, g.C, g.Nout, g.ntaps, s);
  tapgemm_kernel<BN, RESB, TF32><<<grid, kThreads, Cfg::kSmemBytes, s>>>(tmA, tmB, p);
  prof_end(PROF_TAPGEMM, s);
  PP_POST_LAUNCH();
  return PP_OK;
}

// ------------------------------------------------------------------------------------------------
// "pixels on N" variant for layers with <= 128 output channels (stride-1 traversal).
//
// Measured on B200 (profiles/README.md): a 128-row tcgen05.mma costs ~125 cycles + 0.28*N, i.e. narrow N = 64 / 128
// tiles pay 2-3.5x per MAC, and the nine taps of a 3x3 filter re-read the activation tile from L2.  Here the roles
// are swapped and the tap windows share one shared-memory slab:
//   A = weight tile  [128 out-channel rows (rows >= Nout are don't-care) x 64 k]   K-major, 2-D TMA (or resident)
//   B = pixel window [npx = R image rows x Q (<= 256) pixels x 64 k]               K-major
//   D[o, pixel] in TMEM: lane = output channel, column = pixel, two buffers of 256 columns.
// For each horizontal tap offset dw ONE tiled 4-D box {64 ch, Q, R + max_dh, 1 image} (zero-filled outside the
// image) is loaded; the vertical taps dh are windows of that slab starting dh*Q rows (a multiple of 1024 B) further
// down, so they cost no extra L2 traffic.  The epilogue has a channel per thread: BN statistics are plain per-thread
// sums, and the tile is transposed through shared memory so that stores are full NHWC rows.
// ------------------------------------------------------------------------------------------------
constexpr int kPxnMaxGroups = 8;
constexpr int kPxnMaxDh = 8;

struct PxnDev {
  int M, P, Q, PQ;
  int base_h, base_w;
  int C, kchunks, Nout;
  int R, npx, tiles_per_img, num_tiles;
  int ngroups;
  int dw_val[kPxnMaxGroups];
  int ndh[kPxnMaxGroups];
  int dh_val[kPxnMaxGroups][kPxnMaxDh];
  int kofs[kPxnMaxGroups][kPxnMaxDh];
  int wslot[kPxnMaxGroups][kPxnMaxDh];  // resident mode: index of the tap's first weight tile
  int halo_rows, slab_bytes;            // rows per slab, bytes per slab (multiple of 1024)
  int wtile_bytes;                      // Nout * 128
  int max_ndh;
  int resident, res_bytes;              // weights resident in smem; bytes of that region (incl. slack)
  int stage_bytes, nstages;
  int out_H, out_W, out_sh, out_sw, out_ph, out_pw, out_identity;
  void* out;
  int out_f32;
  const float* scale;
  const float* shift;
  int relu;
  float* stats_partial;
  const __nv_bfloat16* add_src;   // see TapEpilogue::add_src
  int dbg;
  int m64;   // Nout == 64: issue M=64 MMAs; 1 = accumulator row r in TMEM lane r, 2 = lane 32*(r/16) + r%16
};

constexpr int kPxnMaxStages = 8;

__global__ void __launch_bounds__(kThreads, 1)
pxn_kernel(const __grid_constant__ CUtensorMap tmAct, const __grid_constant__ CUtensorMap tmW,
           const __grid_constant__ PxnDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* res_base = smem;
  uint8_t* stage_base = smem + p.res_bytes;
  uint8_t* tail = stage_base + (size_t)p.nstages * p.stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint64_t* full = bars;
  uint64_t* empty = bars + kPxnMaxStages;
  uint64_t* tfull = bars + 2 * kPxnMaxStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* bres = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nstages = p.nstages;
  const int ksteps = p.ngroups * p.kchunks;   // pipeline steps per tile: one slab each

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmAct);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < nstages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    mbar_init(bres, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (p.resident) {
      if (elect_one()) {
        int ntiles_w = 0;
        for (int g = 0; g < p.ngroups; ++g) ntiles_w += p.ndh[g] * p.kchunks;
        mbar_arrive_expect_tx(bres, (uint32_t)(ntiles_w * p.wtile_bytes));
        for (int g = 0; g < p.ngroups; ++g)
          for (int j = 0; j < p.ndh[g]; ++j)
            for (int kc = 0; kc < p.kchunks; ++kc)
              tma_load_2d(&tmW, bres, res_base + (size_t)(p.wslot[g][j] + kc) * p.wtile_bytes, p.kofs[g][j] + kc * kBK,
                          0);
      }
      __syncwarp();
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int img = tile / p.tiles_per_img;
      const int p0 = (tile - img * p.tiles_per_img) * p.R;
      for (int g = 0; g < p.ngroups; ++g) {
        const int cw = p.base_w + p.dw_val[g];
        const int ch = p.base_h + p0;
        const int nd = p.ndh[g];
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1, 800 + stage);
          if (elect_one()) {
            uint8_t* slab = stage_base + (size_t)stage * p.stage_bytes;
            const uint32_t bytes = (uint32_t)p.slab_bytes + (p.resident ? 0u : (uint32_t)(nd * p.wtile_bytes));
            if (p.dbg & 16) {
              mbar_arrive(&full[stage]);
            } else {
              mbar_arrive_expect_tx(&full[stage], bytes);
              tma_load_4d(&tmAct, &full[stage], slab, kc * kBK, cw, ch, img);
              if (!p.resident) {
                for (int j = 0; j < nd; ++j)
                  tma_load_2d(&tmW, &full[stage], slab + p.slab_bytes + (size_t)j * p.wtile_bytes,
                              p.kofs[g][j] + kc * kBK, 0);
              }
            }
          }
          __syncwarp();
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc_bf16(p.m64 ? 64 : 128, p.npx, 0, 0);
    const uint32_t stage0 = smem_u32(stage_base);
    const uint32_t res0 = smem_u32(res_base);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    if (p.resident) mbar_wait(bres, 0, 850);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1, 900 + acc);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      int step = 0;
      for (int g = 0; g < p.ngroups; ++g) {
        const int nd = p.ndh[g];
        for (int kc = 0; kc < p.kchunks; ++kc, ++step) {
          mbar_wait(&full[stage], phase, 1000 + stage);
          tc_fence_after();
          const uint32_t slab = stage0 + stage * p.stage_bytes;
          if (elect_one()) {
            for (int j = 0; j < nd; ++j) {
              const uint32_t sw = p.resident ? res0 + (p.wslot[g][j] + kc) * p.wtile_bytes
                                             : slab + p.slab_bytes + j * p.wtile_bytes;
              const uint32_t sx = slab + p.dh_val[g][j] * p.Q * 128;
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k) {
                const uint64_t da = make_smem_desc_sw128(sw + k * 32, 16, 1024);
                const uint64_t db = make_smem_desc_sw128(sx + k * 32, 16, 1024);
                if (!(p.dbg & 8)) tc_mma_bf16(d_tmem, da, db, idesc, (step | j | k) != 0 ? 1u : 0u);
              }
            }
            tc_commit(&empty[stage]);
            if (step == ksteps - 1) tc_commit(&tfull[acc]);
          }
          __syncwarp();
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue: one output channel per thread =====================
    const int quarter = warp & 3;
    int o = quarter * 32 + lane;
    bool o_valid = o < p.Nout;
    if (p.m64 == 2) {   // M=64 accumulator: row r lives in lane 32*(r/16) + r%16
      o = quarter * 16 + (lane & 15);
      o_valid = lane < 16;
    }
    const float ca = (o_valid && p.scale) ? __ldg(p.scale + o) : 1.0f;
    const float cb = (o_valid && p.shift) ? __ldg(p.shift + o) : 0.0f;
    const bool has_affine = (p.scale != nullptr) || (p.shift != nullptr);
    float sum1 = 0.0f, sum2 = 0.0f;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int nchunks = p.npx / 32;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int img = tile / p.tiles_per_img;
      const int p0 = (tile - img * p.tiles_per_img) * p.R;
      const int m0 = img * p.PQ + p0 * p.Q;      // first output pixel (traversal order) of this tile
      mbar_wait(&tfull[acc], acc_phase, 1100 + acc);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + acc * 256;
#pragma unroll 1
      for (int j = 0; j < nchunks; ++j) {
        uint32_t raw[32];
        if (p.dbg & 4) {
#pragma unroll
          for (int i = 0; i < 32; ++i) raw[i] = 0x3f800000u;
        } else {
          tmem_ld_32x32(taddr + j * 32, raw);
          tmem_ld_wait();
        }
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
        if (p.stats_partial && !(p.dbg & 2)) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {   // pixels past the batch tail were zero-filled: they add nothing
            sum1 += v[i];
            sum2 = fmaf(v[i], v[i], sum2);
          }
        }
        if (has_affine) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], ca, cb);
        }
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
        }
        // A lane owns an output channel, so for a fixed pixel the 32 lanes of a warp write 32 consecutive channels:
        // every store instruction is one full 128-byte (fp32) / 64-byte (bf16) NHWC segment — no staging needed.
        // The 32 pixels of a chunk are `32/seg` runs of `seg` consecutive pixels of one image row; inside a run the
        // output row advances by out_sw, so addresses are formed by pointer increments (warp-uniform, no shuffles).
        if (o_valid && !(p.dbg & 1)) {
          const int mc = m0 + j * 32;
          const size_t step = (size_t)(p.out_identity ? 1 : p.out_sw) * p.Nout;
          // Q is a multiple of 8, so each group of 8 consecutive pixels lies in one image row: four row lookups per
          // chunk, pointer increments inside a group.
          size_t base8[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int m = mc + 8 * k;
            size_t orow;
            if (p.out_identity) {
              orow = (size_t)m;
            } else {
              const int rem = m - img * p.PQ;
              const int pp_ = rem / p.Q;
              const int qq_ = rem - pp_ * p.Q;
              orow = ((size_t)img * p.out_H + (size_t)(pp_ * p.out_sh + p.out_ph)) * p.out_W +
                     (size_t)(qq_ * p.out_sw + p.out_pw);
            }
            base8[k] = orow * p.Nout + o;
          }
          if (p.out_f32) {
            float* out = reinterpret_cast<float*>(p.out);
#pragma unroll
            for (int i = 0; i < 32; ++i) out[base8[i >> 3] + (size_t)(i & 7) * step] = v[i];
          } else {
            __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
            if (p.add_src) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const size_t idx = base8[i >> 3] + (size_t)(i & 7) * step;
                const float r = __bfloat162float(__float2bfloat16_rn(v[i])) + __bfloat162float(p.add_src[idx]);
                out[idx] = __float2bfloat16_rn(r);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) out[base8[i >> 3] + (size_t)(i & 7) * step] = __float2bfloat16_rn(v[i]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.stats_partial && o_valid) {
      float* dst = p.stats_partial + (size_t)blockIdx.x * 2 * p.Nout;
      dst[o] = sum1;
      dst[p.Nout + o] = sum2;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

struct PxnPlan {
  bool ok;
  PxnDev dev;
  int grid;
  int smem_bytes;
};

static bool pxn_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PP_NO_PXN"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}

// Decide whether the pixels-on-N kernel applies and fill its launch parameters.
static PxnPlan plan_pxn(const TapGemm& g) {
  PxnPlan pl;
  pl.ok = false;
  if (!pxn_enabled() || g.tf32) return pl;     // TF32 layers (AlexNet: 192 / 256 / 384 channels) take tapgemm_kernel
  if (g.step_h != 1 || g.step_w != 1) return pl;
  if (g.C % 64 != 0 || (g.Nout != 64 && g.Nout != 128)) return pl;
  if (g.Q > 256 || g.Q < 1) return pl;
  int R = 256 / g.Q;
  if (R > g.P) R = g.P;
  while (R > 1 && (g.P % R != 0 || (R * g.Q) % 32 != 0)) --R;
  const int npx = R * g.Q;
  // N = R image rows x Q pixels: 256 on the 32x32 / 16x16 CIFAR maps, 224 (4 rows) on the 56x56 ImageNet maps — a
  // multiple of 32 (epilogue chunk) that keeps the instruction near full width
  if (g.P % R != 0 || npx % 32 != 0 || npx < 192 || R < 2) return pl;
  if (g.Q * 128 % 1024 != 0) return pl;   // tap windows must start on a swizzle-atom boundary: Q % 8 == 0
  PxnDev& d = pl.dev;
  memset(&d, 0, sizeof(d));
  d.M = g.N * g.P * g.Q; d.P = g.P; d.Q = g.Q; d.PQ = g.P * g.Q;
  d.base_h = g.base_h; d.base_w = g.base_w;
  d.C = g.C; d.kchunks = g.C / 64; d.Nout = g.Nout;
  d.R = R; d.npx = npx; d.tiles_per_img = g.P / R; d.num_tiles = g.N * d.tiles_per_img;
  // group the taps by horizontal offset
  int max_dh = 0, wtiles = 0;
  for (int t = 0; t < g.ntaps; ++t) {
    int gi = -1;
    for (int k = 0; k < d.ngroups; ++k)
      if (d.dw_val[k] == g.tap_dw[t]) gi = k;
    if (gi < 0) {
      if (d.ngroups == kPxnMaxGroups) return pl;
      gi = d.ngroups++;
      d.dw_val[gi] = g.tap_dw[t];
    }
    if (d.ndh[gi] == kPxnMaxDh) return pl;
    const int j = d.ndh[gi]++;
    d.dh_val[gi][j] = g.tap_dh[t];
    d.kofs[gi][j] = g.tap_kofs[t];
    d.wslot[gi][j] = wtiles;
    wtiles += d.kchunks;
    if (g.tap_dh[t] > max_dh) max_dh = g.tap_dh[t];
    if (d.ndh[gi] > d.max_ndh) d.max_ndh = d.ndh[gi];
  }
  d.halo_rows = R + max_dh;
  if (d.halo_rows > 256) return pl;
  d.slab_bytes = d.halo_rows * g.Q * 128;
  if (d.slab_bytes % 1024 != 0) return pl;
  d.wtile_bytes = g.Nout * 128;
  const int tail = 256;   // mbarriers + TMEM slot
  const int budget = 232448 - 1024 - tail;
  // resident weights: all (tap, chunk) tiles + one tile of slack (an M=128 MMA reads 128 rows even when Nout = 64)
  const int res_bytes = (wtiles + 1) * d.wtile_bytes + (g.Nout == 64 ? d.wtile_bytes : 0);
  const int res_aligned = (res_bytes + 1023) / 1024 * 1024;
  if (res_aligned + 2 * d.slab_bytes <= budget && res_aligned <= 96 * 1024) {
    d.resident = 1;
    d.res_bytes = res_aligned;
    d.stage_bytes = d.slab_bytes;
  } else {
    d.resident = 0;
    d.res_bytes = 0;
    // slab + weight tiles of the group (+ slack tile for the 128-row read when Nout = 64)
    d.stage_bytes = d.slab_bytes + (d.max_ndh + (g.Nout == 64 ? 1 : 0)) * d.wtile_bytes;
  }
  d.nstages = (budget - d.res_bytes) / d.stage_bytes;
  if (d.nstages > kPxnMaxStages) d.nstages = kPxnMaxStages;
  if (d.nstages < 2) return pl;
  d.out_H = g.out_H; d.out_W = g.out_W; d.out_sh = g.out_sh; d.out_sw = g.out_sw; d.out_ph = g.out_ph;
  d.out_pw = g.out_pw; d.out_identity = g.out_identity;
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  pl.grid = d.num_tiles < sms ? d.num_tiles : sms;
  pl.smem_bytes = 1024 + d.res_bytes + d.nstages * d.stage_bytes + tail;
  pl.ok = true;
  return pl;
}

static int launch_pxn(const TapGemm& g, PxnPlan& pl, const void* act, const void* B, const TapEpilogue& e,
                      cudaStream_t s) {
  CUtensorMap tmAct, tmW;
  PP_TRY(make_map_tiled4d(&tmAct, act, g, g.Q, pl.dev.halo_rows, 1));
  PP_TRY(make_map_2d(&tmW, B, (uint64_t)g.Nout, (uint64_t)g.Ktot, (uint32_t)g.Nout));
  PxnDev& d = pl.dev;
  d.out = e.out; d.out_f32 = e.out_f32; d.scale = e.scale; d.shift = e.shift; d.relu = e.relu;
  d.stats_partial = e.stats_partial;
  d.add_src = (const __nv_bfloat16*)e.add_src;
  PP_REQUIRE(!e.add_src || !e.out_f32, PP_EUNSUPPORTED, "add_src needs a bf16 output");
  {
    static int dbg = -1, m64 = -1;
    if (dbg < 0) { const char* ev = getenv("PP_DEBUG"); dbg = ev ? atoi(ev) : 0; }
    // 64 output channels: M = 64 instructions (accumulator row r in TMEM lane 32*(r/16) + r%16) instead of M = 128 with
    // the upper 64 rows multiplying don't-care data.  Per launch the two cost the same time, but the chip runs this
    // workload at its power cap and the narrower instruction draws less: +0.9 % on the whole step (1905 -> 1953 MHz).
    // PP_M64=0 restores M = 128.
    if (m64 < 0) { const char* ev = getenv("PP_M64"); m64 = ev ? atoi(ev) : 2; }
    d.dbg = dbg;
    d.m64 = (g.Nout == 64) ? m64 : 0;
  }
  PP_SET_MAX_SMEM_ONCE((pxn_kernel), 232448);
  prof_begin(PROF_TAPGEMM, 2.0 * (double)d.M * g.Nout * g.ntaps * g.C, g.C, g.Nout, g.ntaps, s);
  pxn_kernel<<<pl.grid, kThreads, pl.smem_bytes, s>>>(tmAct, tmW, d);
  prof_end(PROF_TAPGEMM, s);
  PP_POST_LAUNCH();
  return PP_OK;
}

int tapgemm_tcgen05_grid(const TapGemm& g) {
  {
    PxnPlan pl = plan_pxn(g);
    if (pl.ok) return pl.grid;
  }
  const int bn = pick_bn(g);
  const long long M = (long long)g.N * g.P * g.Q;
  const long long tiles = ((M + kBM - 1) / kBM) * (g.Nout / bn);
  const int sms = device_sm_count();
  return (int)(tiles < sms ? tiles : sms);
}

int tapgemm_tcgen05(const TapGemm& g, const void* act, const void* B, const TapEpilogue& e, cudaStream_t s) {
  PP_TRY(resolve_encoders());
  PP_REQUIRE(tapgemm_tcgen05_supported(g), PP_EUNSUPPORTED,
             "tcgen05 tap-GEMM needs C%%64==0 and Nout%%64==0 (C=%d Nout=%d)", g.C, g.Nout);
  {
    PxnPlan pl = plan_pxn(g);
    if (pl.ok) return launch_pxn(g, pl, act, B, e, s);
  }
  if (g.tf32) {
    switch (pick_bn(g)) {
      case 256: return launch_tapgemm<256, false, true>(g, act, B, e, s);
      case 192: return launch_tapgemm<192, false, true>(g, act, B, e, s);
      case 128: return launch_tapgemm<128, false, true>(g, act, B, e, s);
      default: return launch_tapgemm<64, false, true>(g, act, B, e, s);
    }
  }
  switch (pick_bn(g)) {
    case 256: return launch_tapgemm<256, false>(g, act, B, e, s);
    case 192: return launch_tapgemm<192, false>(g, act, B, e, s);
    case 128: return launch_tapgemm<128, false>(g, act, B, e, s);
    default:
      if (g.Nout == 64 && g.ntaps * (g.C / kBK) <= kResMaxSteps) return launch_tapgemm<64, true>(g, act, B, e, s);
      return launch_tapgemm<64, false>(g, act, B, e, s);
  }
}

// ------------------------------------------------------------------------------------------------
// The passport block as ONE kernel (models/layers/passportconv2d.py:209-223 + :142-175, sign_loss.py:32-54):
//
//   gamma, beta = W . pool(skey), W . pool(key)       (fp32 master weight, fp64 accumulate; or given: public scale/bias)
//   z = conv(x, W)                                    tcgen05, accumulators stay RESIDENT in TMEM
//   mean, var over the batch                          per-CTA column sums -> global partials -> grid barrier
//   y = relu(gamma * (z - mean) * invstd + beta)      straight from TMEM, bf16, 128-bit stores
//   sign loss / sign accuracy, running statistics
//
// Batch-norm needs the statistics of ALL tiles before any output element exists, so a one-kernel block has to keep
// every accumulator on chip until a grid-wide reduction has happened.  TMEM holds 512 fp32 columns per SM = two
// 128 x 256 tiles; with one CTA per SM that is 2 * #SMs tiles (296 on B200): the passport layers of ResNet-18
// (layer4: 512 output channels, 4x4 / 7x7 maps) fit up to 1184 CIFAR / 386 ImageNet images per GPU.  Launched
// cooperatively (all CTAs co-resident), grid barrier = one atomic counter in the workspace.
//
//   warp 0      TMA producer (as tapgemm_kernel)
//   warp 1      MMA issuer; local tile s accumulates into TMEM columns [256 s, 256 s + 256)
//   warps 2..9  (a) gamma / beta rows of this CTA while the first tile is being computed,
//               (b) pass 1 per tile: z (fp32, saved for backward) out, column sums of z and z^2
//                   (warps 2..5 for a tile whose successor's main loop hides it; all eight — two per TMEM lane
//                   quarter, four 32-column chunks each — for the last tile, which nothing hides),
//               ---- grid barrier ----
//               (c) all warps: fixed-order reduction of the partials -> a = gamma * invstd, b = beta - a * mean,
//               (d) pass 2 per tile: y = relu(a z + b) from TMEM, eight warps.
// The z write of the first tile overlaps the second tile's main loop; after the barrier only y (2 bytes / element) is
// written.  Against the multi-kernel path this removes the re-read of z (4 bytes / element), two launches
// (bn_finalize, affine_apply) and, on the passport path, two more (gemv, sign loss).
// ------------------------------------------------------------------------------------------------
struct FusedDev {
  TapGemmDev g;                 // conv geometry (epilogue fields unused)
  __nv_bfloat16* y;             // [M, Nout] bf16
  float* z;                     // [M, Nout] fp32 (saved for backward)
  float* partial;               // [grid][2][256] per-CTA column sums
  unsigned int* barrier;        // zeroed before the launch
  const float* gamma_in;        // per-channel scale / bias given by the caller (public path, ConvBlock) ...
  const float* beta_in;
  const float* w_oihw;          // ... or derived here from the passport: fp32 master weight [O, C*T],
  const double* Ss;             //     pooled skey / key patches [T*C]
  const double* Sk;
  int Cin, T;
  float* gamma_out;             // [O] (written when derived; global because every CTA needs all of its columns)
  float* beta_out;
  const float* b_sign;          // SignLoss (NULL: none)
  float alpha;
  float* sign_loss;
  float* sign_acc;
  float* rmean; float* rvar;    // running statistics (may be NULL)
  float* save_mean; float* save_invstd;
  float eps, momentum;
  int relu;
  int n_tiles;                  // Nout / 256
};

constexpr int kFusedBN = 256;
constexpr int kFusedThreads = 320;     // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int kFusedStages = 4;
constexpr int kFusedStageBytes = kBM * kBK * 2 + kFusedBN * kBK * 2;   // 48 KiB
constexpr int kFusedSmemBytes = 1024 + kFusedStages * kFusedStageBytes + 2 * kFusedBN * 4 /*a, b*/ +
                                4 * 2 * kFusedBN * 4 /*s_red*/ + 2 * kFusedBN * 4 /*s_acc*/ + 4 * 4096 /*staging*/ + 256;

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// All CTAs of a cooperative launch.  Bounded like every other wait in this file.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int expected) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const long long t0 = clock64();
    while (ld_acquire_u32(counter) < expected) {
      if (clock64() - t0 > 4000000000LL) {
        g_pp_timeout_code = 1500;
        __threadfence_system();
        printf("[passport_sm100] grid barrier timeout block=%d\n", (int)blockIdx.x);
        __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kFusedThreads, 1)
passport_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ FusedDev p) {
  constexpr int BN = kFusedBN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  float* s_a = reinterpret_cast<float*>(stage_base + kFusedStages * kFusedStageBytes);
  float* s_b = s_a + BN;
  float* s_red = s_b + BN;              // [4 warps][2][BN]
  float* s_acc = s_red + 4 * 2 * BN;    // [2][BN] running per-CTA column sums
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_acc + 2 * BN);   // [4 warps][4 KiB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + 4 * 4096);
  uint64_t* full = bars;
  uint64_t* empty = bars + kFusedStages;
  uint64_t* tfull = bars + 2 * kFusedStages;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 2);

  const TapGemmDev& g = p.g;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = g.num_m_tiles * g.num_n_tiles;
  const int kchunks = g.C / kBK;
  const int ksteps = g.ntaps * kchunks;
  const int grid = (int)gridDim.x;
  const int nlocal = ((int)blockIdx.x + grid < num_tiles) ? 2 : 1;     // grid <= num_tiles <= 2 * grid
  const int my_n_tile = (int)blockIdx.x % g.num_n_tiles;               // grid % num_n_tiles == 0: same for both tiles
  const int n0 = my_n_tile * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kFusedStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    const int a_tiled = g.a_tiled;
    for (int s = 0; s < nlocal; ++s) {
      const int tile = (int)blockIdx.x + s * grid;
      const int m_tile = tile / g.num_n_tiles;
      const int m0 = m_tile * kBM;
      const int img = m0 / g.PQ;
      const int rem = m0 - img * g.PQ;
      const int p0 = rem / g.Q;
      const int q0 = rem - p0 * g.Q;
      const int cw = g.base_w + q0 * g.step_w;
      const int ch = g.base_h + p0 * g.step_h;
      for (int t = 0; t < g.ntaps; ++t) {
        const int dw = g.tap_dw[t];
        const int dh = g.tap_dh[t];
        const int kofs = g.tap_kofs[t];
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1, 1600 + stage);
          if (elect_one()) {
            uint8_t* sa = stage_base + stage * kFusedStageBytes;
            uint8_t* sb = sa + kBM * kBK * 2;
            mbar_arrive_expect_tx(&full[stage], kFusedStageBytes);
            if (a_tiled) tma_load_4d(&tmA, &full[stage], sa, kc * kBK, cw + dw, ch + dh, img);
            else tma_load_im2col_4d(&tmA, &full[stage], sa, kc * kBK, cw, ch, img, (uint16_t)dw, (uint16_t)dh);
            tma_load_2d(&tmB, &full[stage], sb, kofs + kc * kBK, n0);
          }
          __syncwarp();
          if (++stage == kFusedStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: tile s -> TMEM columns [256 s, 256 s + 256), never recycled =====================
    constexpr uint32_t idesc = make_idesc_bf16(kBM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t stage0 = smem_u32(stage_base);
    for (int s = 0; s < nlocal; ++s) {
      const uint32_t d_tmem = tmem_base + s * BN;
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full[stage], phase, 1700 + stage);
        tc_fence_after();
        const uint32_t sa = stage0 + stage * kFusedStageBytes;
        const uint32_t sb = sa + kBM * kBK * 2;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t da = make_smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(sb + k * 32, 16, 1024);
            tc_mma_bf16(d_tmem, da, db, idesc, (ks | k) != 0 ? 1u : 0u);
          }
          tc_commit(&empty[stage]);
          if (ks == ksteps - 1) tc_commit(&tfull[s]);
        }
        __syncwarp();
        if (++stage == kFusedStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps, part 1 =====================
    const int ew = warp - 2;              // 0..7
    const int half = ew >> 2;             // 0: warps 2..5, 1: warps 6..9 (join for the last tile and for pass 2)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int tid_e = ew * 32 + lane;     // 0..255
    // per-warp store staging: warps 2..5 own s_stage; warps 6..9 only work once the main loop is over and borrow
    // idle pipeline-stage memory (past the 20 KiB the fp64 scratch of step (c) uses)
    uint8_t* my_stage = half == 0 ? s_stage + (ew & 3) * 4096 : stage_base + 32768 + (ew & 3) * 4096;
    // (a) passport-derived gamma / beta for the channels this CTA owns (8 warps x 1 channel, strided over the grid);
    //     hidden behind the first tile's main loop
    if (p.w_oihw != nullptr) {
      const int K = p.Cin * p.T;                    // one OIHW row; the pooled patches use the same element order
      const int n4 = K >> 2;                        // K % 4 == 0 (C % 64 == 0 on this path)
      for (int c = (int)blockIdx.x * 8 + ew; c < g.Nout; c += 8 * grid) {
        const float4* r4 = reinterpret_cast<const float4*>(p.w_oihw + (size_t)c * K);
        double g4[4] = {0.0, 0.0, 0.0, 0.0}, b4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 4
        for (int i = lane; i < n4; i += 32) {
          const float4 w = __ldg(r4 + i);
          const double2 s0 = *reinterpret_cast<const double2*>(p.Ss + 4 * i);
          const double2 s1 = *reinterpret_cast<const double2*>(p.Ss + 4 * i + 2);
          const double2 k0 = *reinterpret_cast<const double2*>(p.Sk + 4 * i);
          const double2 k1 = *reinterpret_cast<const double2*>(p.Sk + 4 * i + 2);
          g4[0] = fma((double)w.x, s0.x, g4[0]);
          g4[1] = fma((double)w.y, s0.y, g4[1]);
          g4[2] = fma((double)w.z, s1.x, g4[2]);
          g4[3] = fma((double)w.w, s1.y, g4[3]);
          b4[0] = fma((double)w.x, k0.x, b4[0]);
          b4[1] = fma((double)w.y, k0.y, b4[1]);
          b4[2] = fma((double)w.z, k1.x, b4[2]);
          b4[3] = fma((double)w.w, k1.y, b4[3]);
        }
        double gsum = (g4[0] + g4[1]) + (g4[2] + g4[3]);      // same order as pointwise.cu:passport_row_dot, so the
        double bsum = (b4[0] + b4[1]) + (b4[2] + b4[3]);      // bits equal those of the stand-alone GEMV kernel
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
          gsum += __shfl_xor_sync(0xffffffffu, gsum, off);
          bsum += __shfl_xor_sync(0xffffffffu, bsum, off);
        }
        if (lane == 0) {
          p.gamma_out[c] = (float)gsum;
          p.beta_out[c] = (float)bsum;
        }
      }
    }
    if (half == 0)
      for (int i = tid_e; i < 2 * BN; i += 128) s_acc[i] = 0.0f;
    // (b) pass 1: z out (fp32), column sums
    for (int s = 0; s < nlocal; ++s) {
      const bool all8 = (s == nlocal - 1);     // nothing hides the last tile's pass: every epilogue warp takes part
      if (!all8 && half == 1) continue;
      const int jbeg = all8 ? half * 4 : 0;
      const int jend = all8 ? jbeg + 4 : BN / 32;
      const int tile = (int)blockIdx.x + s * grid;
      const int m_tile = tile / g.num_n_tiles;
      const int m = m_tile * kBM + row;
      const bool valid = m < g.M;
      const uint32_t valid_mask = __ballot_sync(0xffffffffu, valid);
      const long long out_row = (long long)m;
      // s_acc zeroed / previous tile's s_red consumed
      if (all8) named_bar_sync(2, 256); else named_bar_sync(1, 128);
      mbar_wait(&tfull[s], 0, 1800 + s);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + s * BN;
#pragma unroll 1
      for (int j = jbeg; j < jend; ++j) {
        uint32_t raw[32];
        tmem_ld_32x32(taddr + j * 32, raw);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = valid ? __uint_as_float(raw[i]) : 0.0f;
        {
          float t1[32], t2[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            t1[i] = v[i];
            t2[i] = v[i] * v[i];
          }
          const float s1 = warp_column_sum(t1, lane);
          const float s2 = warp_column_sum(t2, lane);
          s_red[(quarter * 2 + 0) * BN + j * 32 + lane] = s1;     // one slot per lane quarter: the two warps of a
          s_red[(quarter * 2 + 1) * BN + j * 32 + lane] = s2;     // quarter own disjoint column chunks
        }
        if (p.z != nullptr) {
          float4* st = reinterpret_cast<float4*>(my_stage);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            st[lane * 8 + (i ^ (lane & 7))] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          __syncwarp();
          float* obase = p.z + n0 + j * 32 + (lane & 7) * 4;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + (lane >> 3);
            const float4 val = st[r * 8 + ((lane & 7) ^ (r & 7))];
            const long long orow = __shfl_sync(0xffffffffu, out_row, r);
            if ((valid_mask >> r) & 1u) *reinterpret_cast<float4*>(obase + (size_t)orow * g.Nout) = val;
          }
          __syncwarp();
        }
      }
      if (all8) named_bar_sync(2, 256); else named_bar_sync(1, 128);
      if (half == 0) {
        for (int i = tid_e; i < BN; i += 128) {
          const float a1 = s_red[(0 * 2 + 0) * BN + i] + s_red[(1 * 2 + 0) * BN + i] + s_red[(2 * 2 + 0) * BN + i] +
                           s_red[(3 * 2 + 0) * BN + i];
          const float a2 = s_red[(0 * 2 + 1) * BN + i] + s_red[(1 * 2 + 1) * BN + i] + s_red[(2 * 2 + 1) * BN + i] +
                           s_red[(3 * 2 + 1) * BN + i];
          s_acc[i] += a1;            // column i is only ever touched by this thread
          s_acc[BN + i] += a2;
        }
      }
    }
    if (half == 0) {
      named_bar_sync(1, 128);
      float* dst = p.partial + (size_t)blockIdx.x * 2 * BN;
      for (int i = tid_e; i < 2 * BN; i += 128) dst[i] = s_acc[i];
    }
  }

  // ===================== grid-wide: every tile's statistics are in global memory =====================
  tc_fence_before();
  grid_barrier(p.barrier, (unsigned int)grid);
  tc_fence_after();

  // (c) all 192 threads: fixed-order reduction of the partial rows of the CTAs that share this CTA's 256 columns.
  //     The pipeline stages are idle now: their memory holds the fp64 scratch.
  {
    constexpr int kGroups = kFusedThreads / 64;             // 5
    double* sh = reinterpret_cast<double*>(stage_base);     // [kGroups][2 * BN] = 20 KiB
    const int grp = threadIdx.x / 64;                       // rows k == grp (mod kGroups)
    const int l64 = threadIdx.x % 64;                       // float4 #l64 of sum z, float4 #l64 of sum z^2
    const int nrows = grid / g.num_n_tiles;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // the loads of different rows are independent (only the fp64 adds are ordered): keep eight rows in flight
#pragma unroll 8
    for (int k = grp; k < nrows; k += kGroups) {
      const float4* rowp = reinterpret_cast<const float4*>(p.partial + (size_t)(my_n_tile + k * g.num_n_tiles) * 2 * BN);
      const float4 u = __ldcg(rowp + l64);
      const float4 w = __ldcg(rowp + 64 + l64);
      acc[0] += u.x; acc[1] += u.y; acc[2] += u.z; acc[3] += u.w;
      acc[4] += w.x; acc[5] += w.y; acc[6] += w.z; acc[7] += w.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sh[grp * 2 * BN + l64 * 4 + i] = acc[i];
      sh[grp * 2 * BN + BN + l64 * 4 + i] = acc[4 + i];
    }
    __syncthreads();
    const double n = (double)g.M;
    for (int col = threadIdx.x; col < BN; col += kFusedThreads) {
      double s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int gq = 0; gq < kGroups; ++gq) {                 // fixed group order
        s1 += sh[gq * 2 * BN + col];
        s2 += sh[gq * 2 * BN + BN + col];
      }
      const double mu = s1 / n;
      double var = s2 / n - mu * mu;
      if (var < 0.0) var = 0.0;
      const float mean = (float)mu;
      const float invstd = (float)(1.0 / sqrt(var + (double)p.eps));
      const int o = n0 + col;
      const float gm = p.w_oihw ? __ldcg(p.gamma_out + o) : (p.gamma_in ? __ldg(p.gamma_in + o) : 1.0f);
      const float bt = p.w_oihw ? __ldcg(p.beta_out + o) : (p.beta_in ? __ldg(p.beta_in + o) : 0.0f);
      const float a = gm * invstd;
      s_a[col] = a;
      s_b[col] = bt - a * mean;
      if ((int)blockIdx.x < g.num_n_tiles) {      // one CTA per column block owns the per-channel side effects
        if (p.save_mean) p.save_mean[o] = mean;
        if (p.save_invstd) p.save_invstd[o] = invstd;
        if (p.rmean) {
          const double unbiased = g.M > 1 ? var * (n / (n - 1.0)) : var;
          p.rmean[o] = (1.0f - p.momentum) * p.rmean[o] + p.momentum * mean;
          p.rvar[o] = (1.0f - p.momentum) * p.rvar[o] + p.momentum * (float)unbiased;
        }
      }
    }
    // SignLoss.add(gamma) (sign_loss.py:25-28, 53-54): the last CTA, fixed-order fp64 tree
    if (p.b_sign != nullptr && (p.sign_loss || p.sign_acc) && (int)blockIdx.x == grid - 1) {
      __syncthreads();                                       // sh is about to be reused
      double h = 0.0, r = 0.0, a = 0.0;
      for (int o = threadIdx.x; o < g.Nout; o += kFusedThreads) {
        const float gm = p.w_oihw ? __ldcg(p.gamma_out + o) : (p.gamma_in ? __ldg(p.gamma_in + o) : 1.0f);
        const float bb = __ldg(p.b_sign + o);
        const float hinge = fmaxf(-bb * gm + 0.1f, 0.0f);
        h += (double)(p.alpha * hinge);
        r += (double)(gm * gm);
        const float sb = (bb > 0.f) - (bb < 0.f);
        const float sg = (gm > 0.f) - (gm < 0.f);
        a += (sb == sg) ? 1.0 : 0.0;
      }
      sh[threadIdx.x] = h;
      sh[kFusedThreads + threadIdx.x] = r;
      sh[2 * kFusedThreads + threadIdx.x] = a;
      __syncthreads();
      if (threadIdx.x == 0) {
        double hs = 0.0, rs = 0.0, as = 0.0;
        for (int i = 0; i < kFusedThreads; ++i) {
          hs += sh[i]; rs += sh[kFusedThreads + i]; as += sh[2 * kFusedThreads + i];
        }
        if (p.sign_loss) *p.sign_loss = (float)(hs + 0.00001 * rs);
        if (p.sign_acc) *p.sign_acc = (float)(as / (double)g.Nout);
      }
    }
    __syncthreads();
  }

  // (d) pass 2: y = relu(a z + b) straight from the resident accumulators — eight warps, four column chunks each
  if (warp >= 2) {
    const int ew = warp - 2;
    const int half = ew >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    uint8_t* my_stage = half == 0 ? s_stage + (ew & 3) * 4096 : stage_base + 32768 + (ew & 3) * 4096;
    for (int s = 0; s < nlocal; ++s) {
      const int tile = (int)blockIdx.x + s * grid;
      const int m_tile = tile / g.num_n_tiles;
      const int m = m_tile * kBM + row;
      const bool valid = m < g.M;
      const uint32_t valid_mask = __ballot_sync(0xffffffffu, valid);
      const long long out_row = (long long)m;
      const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + s * BN;
#pragma unroll 1
      for (int j = half * 4; j < half * 4 + 4; ++j) {
        uint32_t raw[32];
        tmem_ld_32x32(taddr + j * 32, raw);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaf(__uint_as_float(raw[i]), s_a[j * 32 + i], s_b[j * 32 + i]);
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
        }
        uint4* st = reinterpret_cast<uint4*>(my_stage);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
          u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
          u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
          u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
          st[lane * 4 + (i ^ ((lane >> 1) & 3))] = u;
        }
        __syncwarp();
        __nv_bfloat16* obase = p.y + n0 + j * 32 + (lane & 3) * 8;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int r = it * 8 + (lane >> 2);
          const uint4 val = st[r * 4 + ((lane & 3) ^ ((r >> 1) & 3))];
          const long long orow = __shfl_sync(0xffffffffu, out_row, r);
          if ((valid_mask >> r) & 1u) *reinterpret_cast<uint4*>(obase + (size_t)orow * g.Nout) = val;
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int g_fused_on = -1;     // -1: not decided yet (env PP_NO_FUSED=1 disables); pp_debug_fused() overrides
static bool fused_enabled() {
  if (g_fused_on < 0) { const char* e = getenv("PP_NO_FUSED"); g_fused_on = (e && e[0] == '1') ? 0 : 1; }
  return g_fused_on == 1;
}
int debug_fused(int on) {
  const int prev = fused_enabled() ? 1 : 0;
  if (on >= 0) g_fused_on = on ? 1 : 0;
  return prev;
}

// Can the block run as the single cooperative kernel?  (output channels in 256-wide column blocks, an identity
// output mapping, and every tile resident in TMEM: tiles <= 2 * grid with grid a multiple of the column-block count)
static int fused_grid_for(const TapGemm& g) {
  if (!fused_enabled() || g.tf32 || !tapgemm_tcgen05_supported(g)) return 0;
  if (g.Nout % kFusedBN != 0 || !g.out_identity) return 0;
  const long long M = (long long)g.N * g.P * g.Q;
  const long long tiles = ((M + kBM - 1) / kBM) * (g.Nout / kFusedBN);
  int sms = device_sm_count();
  if (sms <= 0) return 0;
  const int n_tiles = g.Nout / kFusedBN;
  int grid = sms - sms % n_tiles;
  if (tiles < grid) grid = (int)(tiles - tiles % n_tiles);
  if (grid < n_tiles || tiles > 2LL * grid) return 0;
  return grid;
}

bool passport_fused_supported(const TapGemm& g) { return fused_grid_for(g) > 0; }
int passport_fused_grid(const TapGemm& g) { return fused_grid_for(g); }

int passport_fused_tcgen05(const TapGemm& g, const void* act, const void* B, const FusedArgs& a, cudaStream_t s) {
  PP_TRY(resolve_encoders());
  const int grid = fused_grid_for(g);
  PP_REQUIRE(grid > 0, PP_EUNSUPPORTED, "fused passport kernel: geometry not resident in TMEM (M=%d Nout=%d)",
             g.N * g.P * g.Q, g.Nout);
  CUtensorMap tmA, tmB;
  int bw = 0, bh = 0, bn = 0;
  const bool a_tiled = prefer_tiled() && tiled_box_for(g, kBM, &bw, &bh, &bn);
  if (a_tiled) PP_TRY(make_map_tiled4d(&tmA, act, g, bw, bh, bn));
  else PP_TRY(make_map_im2col(&tmA, act, g, kBM));
  PP_TRY(make_map_2d(&tmB, B, (uint64_t)g.Nout, (uint64_t)g.Ktot, kFusedBN));
  FusedDev p;
  memset(&p, 0, sizeof(p));
  TapGemmDev& d = p.g;
  d.a_tiled = a_tiled ? 1 : 0;
  d.M = g.N * g.P * g.Q;
  d.P = g.P; d.Q = g.Q; d.PQ = g.P * g.Q;
  d.base_h = g.base_h; d.base_w = g.base_w; d.step_h = g.step_h; d.step_w = g.step_w;
  d.C = g.C; d.ntaps = g.ntaps; d.Nout = g.Nout;
  d.num_m_tiles = (d.M + kBM - 1) / kBM;
  d.num_n_tiles = g.Nout / kFusedBN;
  d.out_identity = 1;
  for (int t = 0; t < g.ntaps; ++t) {
    d.tap_dh[t] = g.tap_dh[t]; d.tap_dw[t] = g.tap_dw[t]; d.tap_kofs[t] = g.tap_kofs[t];
  }
  p.y = (__nv_bfloat16*)a.y; p.z = a.z; p.partial = a.partial; p.barrier = a.barrier;
  p.gamma_in = a.gamma_in; p.beta_in = a.beta_in;
  p.w_oihw = a.w_oihw; p.Ss = a.Ss; p.Sk = a.Sk; p.Cin = a.Cin; p.T = a.T;
  p.gamma_out = a.gamma_out; p.beta_out = a.beta_out;
  p.b_sign = a.b_sign; p.alpha = a.alpha; p.sign_loss = a.sign_loss; p.sign_acc = a.sign_acc;
  p.rmean = a.rmean; p.rvar = a.rvar; p.save_mean = a.save_mean; p.save_invstd = a.save_invstd;
  p.eps = a.eps; p.momentum = a.momentum; p.relu = a.relu;
  p.n_tiles = d.num_n_tiles;
  PP_REQUIRE(!p.w_oihw || (p.Ss && p.Sk && p.gamma_out && p.beta_out), PP_EBADARG,
             "fused passport kernel: pooled keys / gamma, beta buffers missing");
  PP_SET_MAX_SMEM_ONCE((passport_fused_kernel), kFusedSmemBytes);
  PP_CHECK_CUDA(cudaMemsetAsync(a.barrier, 0, sizeof(unsigned int), s));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kFusedThreads);
  cfg.dynamicSmemBytes = kFusedSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  prof_begin(PROF_FUSED, 2.0 * (double)d.M * g.Nout * g.ntaps * g.C, g.C, g.Nout, g.ntaps, s);
  const cudaError_t err = cudaLaunchKernelEx(&cfg, passport_fused_kernel, tmA, tmB, p);
  prof_end(PROF_FUSED, s);
  if (err != cudaSuccess) {
    set_error("cooperative launch of passport_fused_kernel failed: %s (grid %d)", cudaGetErrorString(err), grid);
    return PP_ELAUNCH;
  }
  PP_POST_LAUNCH();
  return PP_OK;
}

// ------------------------------------------------------------------------------------------------
// weight gradient, computed transposed:  D[(tap, c), o] = sum over pixels m of x_tap[m, c] * dz[m, o]
//   A = x_tap : MN-major ((tap, c) contiguous per pixel row), two 64-wide slabs per CTA loaded in im2col mode
//               (a slab is 64 channels of ONE tap, so C % 64 == 0 keeps slabs from straddling taps)
//   B = dz    : MN-major (o contiguous), BN/64 slabs of [64 pixels x 64 o]
// Putting (tap, c) on M keeps all 128 MMA rows busy even for 64-channel layers, and BN up to 256 output
// channels per tile cuts the smem bytes read per MAC.  The reduction (pixels) is split over blockIdx.y;
// fp32 partial tiles go to partial[split][o][tap*C + c] and are summed by pointwise.cu:wgrad_finalize.
// ------------------------------------------------------------------------------------------------
constexpr int kWK = 64;  // pixels per stage

// TF32 (fp32 x / dz): MN-major operands of a 4-byte type must use the 128B_BASE32B layout — 128-byte rows (32 channels)
// whose 32-byte chunks are swizzled with the row index mod 4 (TMA SWIZZLE_128B_ATOM_32B), 4-row k groups 512 B apart.
// A slab is then 32 channels wide: four x slabs fill the 128 accumulator rows, BN/32 dz slabs the columns, a stage is
// 32 pixels and one kind::tf32 instruction contracts 8 of them (1024 B further down every slab).
template <int BN, bool TF32 = false>
struct WgCfg {
  static constexpr int kSlabElems = TF32 ? 32 : 64;      // (tap, c) / o entries per 128-byte row
  static constexpr int kPix = TF32 ? 32 : kWK;           // pixels per stage
  static constexpr int kSlabBytes = kPix * 128;
  static constexpr int kSlabsA = 128 / kSlabElems;
  static constexpr int kSlabsB = BN / kSlabElems;
  static constexpr int kStageA = kSlabsA * kSlabBytes;   // 16 KiB
  static constexpr int kStageB = kSlabsB * kSlabBytes;   // 8 / 16 / 32 KiB
  static constexpr int kStageBytes = kStageA + kStageB;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 192 ? 5 : (BN == 128 ? 6 : 8));
  static constexpr int kPixPerMma = TF32 ? 8 : 16;
  static constexpr int kTmemCols = BN == 192 ? 256 : BN;   // allocations are powers of 2
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 256;
  static_assert(kSmemBytes <= 232448, "wgrad stage configuration exceeds the 227 KiB of shared memory per CTA");
};

struct WgradDev {
  int M, P, Q, PQ;
  int base_h, base_w, step_h, step_w;
  int C, O, ntaps, Ktot;   // Ktot = ntaps * C = rows of the transposed gradient
  int n_tiles;             // O / BN
  int x_tiled;             // 1: x slabs are loaded with a tiled 4-D box instead of im2col mode
  int chunks_total, chunks_per_split;
  int8_t tap_dh[kMaxTaps], tap_dw[kMaxTaps];
  float* partial;  // [splits][O][Ktot]
};

template <int BN, bool TF32 = false>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDz,
             const __grid_constant__ WgradDev p) {
  using Cfg = WgCfg<BN, TF32>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::kStages;
  uint64_t* tfull = bars + 2 * Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int n_tile = blockIdx.x % p.n_tiles;
  const int m_tile = blockIdx.x / p.n_tiles;
  const int split = blockIdx.y;
  const int chunk_lo = split * p.chunks_per_split;
  int chunk_hi = chunk_lo + p.chunks_per_split;
  if (chunk_hi > p.chunks_total) chunk_hi = p.chunks_total;
  const int nchunks = chunk_hi > chunk_lo ? chunk_hi - chunk_lo : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDz);
    tma_prefetch_desc(&tmX);
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // TMA producer: warp-uniform loop, one elected lane issues the copies
    // the A slabs of this tile: rows [k_i, k_i + kSlabElems) of the (tap, c) axis, k_i = m_tile*128 + i*kSlabElems
    int c0[Cfg::kSlabsA], dw[Cfg::kSlabsA], dh[Cfg::kSlabsA];
#pragma unroll
    for (int i = 0; i < Cfg::kSlabsA; ++i) {
      int k = m_tile * 128 + i * Cfg::kSlabElems;
      if (k >= p.Ktot) k = m_tile * 128;   // past the end: load slab 0 again (its rows are not stored)
      const int tap = k / p.C;
      c0[i] = k - tap * p.C;
      dw[i] = p.tap_dw[tap];
      dh[i] = p.tap_dh[tap];
    }
    const int x_tiled = p.x_tiled;
    int stage = 0;
    uint32_t phase = 0;
    for (int ch = chunk_lo; ch < chunk_lo + nchunks; ++ch) {
      const int m0 = ch * Cfg::kPix;
      const int img = m0 / p.PQ;
      const int rem = m0 - img * p.PQ;
      const int p0 = rem / p.Q;
      const int q0 = rem - p0 * p.Q;
      const int cw = p.base_w + q0 * p.step_w;
      const int chh = p.base_h + p0 * p.step_h;
      mbar_wait(&empty[stage], phase ^ 1, 500 + stage);
      if (elect_one()) {
        uint8_t* sa = smem + stage * Cfg::kStageBytes;
        uint8_t* sb = sa + Cfg::kStageA;
        mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
#pragma unroll
        for (int i = 0; i < Cfg::kSlabsA; ++i) {
          if (x_tiled) tma_load_4d(&tmX, &full[stage], sa + i * Cfg::kSlabBytes, c0[i], cw + dw[i], chh + dh[i], img);
          else tma_load_im2col_4d(&tmX, &full[stage], sa + i * Cfg::kSlabBytes, c0[i], cw, chh, img, (uint16_t)dw[i],
                                  (uint16_t)dh[i]);
        }
#pragma unroll
        for (int i = 0; i < Cfg::kSlabsB; ++i)
          tma_load_2d(&tmDz, &full[stage], sb + i * Cfg::kSlabBytes, n_tile * BN + i * Cfg::kSlabElems, m0);
      }
      __syncwarp();
      if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // MMA issuer: warp-uniform loop, one elected lane issues tcgen05.mma / commit
    constexpr uint32_t idesc = TF32 ? make_idesc_tf32(128, BN, 1, 1) : make_idesc_bf16(128, BN, 1, 1);
    const uint32_t smem0 = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < nchunks; ++it) {
      mbar_wait(&full[stage], phase, 600 + stage);
      tc_fence_after();
      const uint32_t sa = smem0 + stage * Cfg::kStageBytes;
      const uint32_t sb = sa + Cfg::kStageA;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < Cfg::kPix / Cfg::kPixPerMma; ++k) {
          if (TF32) {
            // 8 pixels (k) per instruction = two 4-row groups of 512 B; MN slabs are kSlabBytes apart
            const uint64_t da = make_smem_desc_mn32(sa + k * 1024, Cfg::kSlabBytes, 512);
            const uint64_t db = make_smem_desc_mn32(sb + k * 1024, Cfg::kSlabBytes, 512);
            tc_mma_tf32(tmem_base, da, db, idesc, (it | k) != 0 ? 1u : 0u);
          } else {
            // 16 pixels (k) per instruction = two 8-row groups of 1024 B; MN slabs are kSlabBytes apart
            const uint64_t da = make_smem_desc_sw128(sa + k * 2048, Cfg::kSlabBytes, 1024);
            const uint64_t db = make_smem_desc_sw128(sb + k * 2048, Cfg::kSlabBytes, 1024);
            tc_mma_bf16(tmem_base, da, db, idesc, (it | k) != 0 ? 1u : 0u);
          }
        }
        tc_commit(&empty[stage]);
        if (it == nchunks - 1) tc_commit(tfull);
      }
      __syncwarp();
      if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int k = m_tile * 128 + row;           // (tap, c) index of this accumulator row
    const bool valid = k < p.Ktot;
    if (nchunks > 0) {
      mbar_wait(tfull, 0, 700);
      tc_fence_after();
    }
    const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16);
    float* dst0 = p.partial + ((size_t)split * p.O + (size_t)n_tile * BN) * p.Ktot + k;
#pragma unroll 1
    for (int j = 0; j < BN / 32; ++j) {
      uint32_t raw[32];
      if (nchunks > 0) {
        tmem_ld_32x32(taddr + j * 32, raw);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) raw[i] = 0u;
      }
      if (valid) {
        // lanes hold consecutive k: every store instruction writes 32 consecutive floats of one o row
        float* dst = dst0 + (size_t)(j * 32) * p.Ktot;
#pragma unroll
        for (int i = 0; i < 32; ++i) dst[(size_t)i * p.Ktot] = __uint_as_float(raw[i]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

bool wgrad_tcgen05_supported(const TapGemm& g, int O) {
  if (g.C % (g.tf32 ? 32 : 64) != 0 || O % 64 != 0) return false;
  if (g.base_h < -128 || g.base_w < -128 || g.upper_h < -128 || g.upper_w < -128) return false;
  if (g.base_h > 127 || g.base_w > 127 || g.upper_h > 127 || g.upper_w > 127) return false;
  for (int t = 0; t < g.ntaps; ++t)
    if (g.tap_dh[t] < 0 || g.tap_dw[t] < 0) return false;
  return true;
}

static int wgrad_bn(int O) {
  if (O % 256 == 0) return 256;
  if (O % 192 == 0) return 192;      // AlexNet: 192 / 384 output channels
  return (O % 128 == 0) ? 128 : 64;
}

static bool wgrad_om_applies(const TapGemm& g, int O);
static bool wgrad_om_paired(const TapGemm& g, int O);

int wgrad_pick_splits(const TapGemm& g, int O) {
  const int Ktot = g.ntaps * g.C;
  const int tiles = wgrad_om_paired(g, O) ? 2
                    : wgrad_om_applies(g, O) ? (Ktot / 64 + 3) / 4 : ((Ktot + 127) / 128) * (O / wgrad_bn(O));
  const long long M = (long long)g.N * g.P * g.Q;
  const int pix = g.tf32 ? 32 : kWK;
  const int chunks = (int)((M + pix - 1) / pix);
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  int splits = (2 * sms) / tiles;          // two full waves when the tile count allows it
  if (splits > chunks / 4) splits = chunks / 4;   // at least a few pipeline stages of work per CTA
  if (splits < 1) splits = 1;
  // keep the fp32 partial workspace bounded (64 MiB)
  const long long per_split = (long long)O * Ktot * 4;
  while (splits > 1 && per_split * splits > (64ll << 20)) --splits;
  return splits;
}

template <int BN, bool TF32 = false>
static int launch_wgrad(const TapGemm& g, const void* x, const void* dz, int O, float* partial, int splits,
                        cudaStream_t s) {
  using Cfg = WgCfg<BN, TF32>;
  CUtensorMap tmDz, tmX;
  const long long M = (long long)g.N * g.P * g.Q;
  PP_TRY(make_map_2d(&tmDz, dz, (uint64_t)M, (uint64_t)O, Cfg::kPix, TF32 ? 4 : 2, TF32));
  int bw = 0, bh = 0, bn = 0;
  const bool x_tiled = !TF32 && prefer_tiled() && tiled_box_for(g, kWK, &bw, &bh, &bn);
  if (x_tiled) PP_TRY(make_map_tiled4d(&tmX, x, g, bw, bh, bn));
  else PP_TRY(make_map_im2col(&tmX, x, g, Cfg::kPix, TF32));
  WgradDev p;
  p.x_tiled = x_tiled ? 1 : 0;
  p.M = (int)M; p.P = g.P; p.Q = g.Q; p.PQ = g.P * g.Q;
  p.base_h = g.base_h; p.base_w = g.base_w; p.step_h = g.step_h; p.step_w = g.step_w;
  p.C = g.C; p.O = O; p.ntaps = g.ntaps; p.Ktot = g.ntaps * g.C;
  p.n_tiles = O / BN;
  p.chunks_total = (int)((M + Cfg::kPix - 1) / Cfg::kPix);
  p.chunks_per_split = (p.chunks_total + splits - 1) / splits;
  for (int t = 0; t < g.ntaps; ++t) { p.tap_dh[t] = g.tap_dh[t]; p.tap_dw[t] = g.tap_dw[t]; }
  p.partial = partial;
  PP_SET_MAX_SMEM_ONCE((wgrad_kernel<BN, TF32>), Cfg::kSmemBytes);
  const int m_tiles = (p.Ktot + 127) / 128;
  dim3 grid(m_tiles * p.n_tiles, splits);
  prof_begin(PROF_WGRAD, 2.0 * (double)M * O * g.ntaps * g.C, g.C, O, g.ntaps, s);
  wgrad_kernel<BN, TF32><<<grid, kThreads, Cfg::kSmemBytes, s>>>(tmX, tmDz, p);
  prof_end(PROF_WGRAD, s);
  PP_POST_LAUNCH();
  return PP_OK;
}

// ------------------------------------------------------------------------------------------------
// weight gradient for O <= 128: output channels on M, (tap, c) on N = 256.
//   D[o, (tap,c)] = sum over pixels of dz[m, o] * x_tap[m, c]
// A 128-row tcgen05.mma costs ~125 cycles + 0.28*N (profiles/README.md), so with only 64 / 128 output channels the
// transposed kernel above (N = O) pays 142 / 155 cycles for a quarter / half of the MACs an N = 256 instruction does
// in 197.  Here A = dz slabs (MN-major, o contiguous; the second slab is TMA-zero-filled when O = 64) and
// B = four 64-wide (tap, c) slabs of x.
// ------------------------------------------------------------------------------------------------
//
// Tap pairing (O == C == 64, 3x3 / stride 1 / pad 1: the layer1 convs).  With O = 64 the upper 64 accumulator rows
// would multiply zeros.  Instead they get dz shifted down by one image row (a 4-D TMA box on dz, rows past the image
// zero-filled): against the x slab of tap (dh, dw) the lower rows accumulate dW[(dh, dw)] and the upper rows
//   sum_pixel dz[pixel + Q] x[pixel + (dh, dw)] = sum_pixel' dz[pixel'] x[pixel' + (dh - 1, dw)] = dW[(dh - 1, dw)]
// (the pixel' in image row 0 that the shift drops only ever meets the zero padding row when dh - 1 = 0).  Two CTAs
// columns of N = 192 — taps (1, .) giving dW[1, .] and dW[0, .], taps (2, .) giving dW[2, .] — replace three of
// N = 256: 6 slab products instead of 12 for the 9 useful ones.
struct WgradOmDev {
  int M, P, Q, PQ;
  int base_h, base_w, step_h, step_w;
  int C, O, ntaps, Ktot, nslabs;
  int x_tiled;
  int paired;   // 1: tap pairing (see above); grid.x == 2
  int chunks_total, chunks_per_split;
  int8_t tap_dh[kMaxTaps], tap_dw[kMaxTaps];
  float* partial;  // [splits][O][Ktot]
};

constexpr int kOmStageA = 2 * kWK * 128;   // two dz slabs (o 0..63, 64..127)
constexpr int kOmStageB = 4 * kWK * 128;   // four (tap, c) slabs of x
constexpr int kOmStageBytes = kOmStageA + kOmStageB;   // 48 KiB
constexpr int kOmStages = 4;
constexpr int kOmSmemBytes = 1024 + kOmStages * kOmStageBytes + 256;

__global__ void __launch_bounds__(kThreads, 1)
wgrad_om_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDz,
                const __grid_constant__ CUtensorMap tmDzS /*dz as [N,P,Q,O], row-shifted loads (paired mode)*/,
                const __grid_constant__ WgradOmDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOmStages * kOmStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kOmStages;
  uint64_t* tfull = bars + 2 * kOmStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x;            // four consecutive (tap, c) slabs
  const int split = blockIdx.y;
  const int chunk_lo = split * p.chunks_per_split;
  int chunk_hi = chunk_lo + p.chunks_per_split;
  if (chunk_hi > p.chunks_total) chunk_hi = p.chunks_total;
  const int nchunks = chunk_hi > chunk_lo ? chunk_hi - chunk_lo : 0;
  const int paired = p.paired;
  const int nb = paired ? 3 : 4;            // x slabs per stage

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDz);
    tma_prefetch_desc(&tmX);
    if (paired) tma_prefetch_desc(&tmDzS);
    for (int i = 0; i < kOmStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    int tap[4], c0[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int sidx = n_tile * 4 + i;
      if (sidx >= p.nslabs) sidx = n_tile * 4;   // past the end: reload the first slab (its columns are not stored)
      const int k0 = sidx * 64;
      tap[i] = k0 / p.C;
      c0[i] = k0 - tap[i] * p.C;
      if (paired) {   // C == 64: slab == tap; this CTA column owns filter row dh = n_tile + 1
        tap[i] = (n_tile + 1) * 3 + (i < 3 ? i : 0);
        c0[i] = 0;
      }
    }
    const int x_tiled = p.x_tiled;
    const uint32_t stage_tx = kOmStageA + nb * kWK * 128;
    int stage = 0;
    uint32_t phase = 0;
    for (int ch = chunk_lo; ch < chunk_lo + nchunks; ++ch) {
      const int m0 = ch * kWK;
      const int img = m0 / p.PQ;
      const int rem = m0 - img * p.PQ;
      const int p0 = rem / p.Q;
      const int q0 = rem - p0 * p.Q;
      const int cw = p.base_w + q0 * p.step_w;
      const int chh = p.base_h + p0 * p.step_h;
      mbar_wait(&empty[stage], phase ^ 1, 1200 + stage);
      if (elect_one()) {
        uint8_t* sa = smem + stage * kOmStageBytes;
        uint8_t* sb = sa + kOmStageA;
        mbar_arrive_expect_tx(&full[stage], stage_tx);
        tma_load_2d(&tmDz, &full[stage], sa, 0, m0);
        if (paired && n_tile == 0)   // dz one image row further down; rows past the image read as zeros
          tma_load_4d(&tmDzS, &full[stage], sa + kWK * 128, 0, 0, p0 + 1, img);
        else
          tma_load_2d(&tmDz, &full[stage], sa + kWK * 128, 64, m0);   // all zeros (out of bounds) when O == 64
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i >= nb) break;
          const int dw = p.tap_dw[tap[i]], dh = p.tap_dh[tap[i]];
          if (x_tiled) tma_load_4d(&tmX, &full[stage], sb + i * kWK * 128, c0[i], cw + dw, chh + dh, img);
          else tma_load_im2col_4d(&tmX, &full[stage], sb + i * kWK * 128, c0[i], cw, chh, img, (uint16_t)dw,
                                  (uint16_t)dh);
        }
      }
      __syncwarp();
      if (++stage == kOmStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = paired ? make_idesc_bf16(128, 192, 1, 1) : make_idesc_bf16(128, 256, 1, 1);
    const uint32_t smem0 = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < nchunks; ++it) {
      mbar_wait(&full[stage], phase, 1300 + stage);
      tc_fence_after();
      const uint32_t sa = smem0 + stage * kOmStageBytes;
      const uint32_t sb = sa + kOmStageA;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kWK / 16; ++k) {
          const uint64_t da = make_smem_desc_sw128(sa + k * 2048, kWK * 128, 1024);
          const uint64_t db = make_smem_desc_sw128(sb + k * 2048, kWK * 128, 1024);
          tc_mma_bf16(tmem_base, da, db, idesc, (it | k) != 0 ? 1u : 0u);
        }
        tc_commit(&empty[stage]);
        if (it == nchunks - 1) tc_commit(tfull);
      }
      __syncwarp();
      if (++stage == kOmStages) { stage = 0; phase ^= 1; }
    }
  } else {
    const int quarter = warp & 3;
    const int o = quarter * 32 + lane;
    if (nchunks > 0) {
      mbar_wait(tfull, 0, 1400);
      tc_fence_after();
    }
    const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16);
    float* dst_row = p.partial + ((size_t)split * p.O + (size_t)o) * p.Ktot + (size_t)n_tile * 256;
    int nj = 8;
    bool row_valid = o < p.O;
    if (paired) {
      // rows 0..63: dW[o][(dh = n_tile + 1, dw)] ; rows 64..127: dW[o - 64][(dh - 1, dw)], kept for dh - 1 == 0 only
      const bool hi = o >= 64;
      const int orow = hi ? o - 64 : o;
      const int kbase = hi ? 0 : (n_tile + 1) * 192;
      row_valid = !hi || n_tile == 0;
      dst_row = p.partial + ((size_t)split * p.O + (size_t)orow) * p.Ktot + kbase;
      nj = 6;
    }
#pragma unroll 1
    for (int j = 0; j < nj; ++j) {
      uint32_t raw[32];
      if (nchunks > 0) {
        tmem_ld_32x32(taddr + j * 32, raw);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) raw[i] = 0u;
      }
      const bool slab_valid = paired || (n_tile * 4 + (j >> 1)) < p.nslabs;
      if (row_valid && slab_valid) {
        float4* dst = reinterpret_cast<float4*>(dst_row + j * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          dst[i] = make_float4(__uint_as_float(raw[4 * i]), __uint_as_float(raw[4 * i + 1]),
                               __uint_as_float(raw[4 * i + 2]), __uint_as_float(raw[4 * i + 3]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

static bool wgrad_om_applies(const TapGemm& g, int O) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("PP_NO_WGRAD_OM"); off = (e && e[0] == '1') ? 1 : 0; }
  return !off && !g.tf32 && (O == 64 || O == 128) && g.C % 64 == 0;
}

// tap pairing needs: O == C == 64, canonical 3x3 / stride 1 / pad 1 tap list, and chunks of kWK pixels that are whole
// rows of one image (so the row-shifted dz chunk is a 4-D box)
static bool wgrad_om_paired(const TapGemm& g, int O) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("PP_NO_WGRAD_PAIR"); off = (e && e[0] == '1') ? 1 : 0; }
  if (off || !wgrad_om_applies(g, O)) return false;
  if (O != 64 || g.C != 64 || g.ntaps != 9 || g.step_h != 1 || g.step_w != 1 || g.base_h != -1 || g.base_w != -1)
    return false;
  if (g.P != g.H || g.Q != g.W || g.Q > kWK || kWK % g.Q != 0 || (g.P * g.Q) % kWK != 0) return false;
  for (int t = 0; t < 9; ++t)
    if (g.tap_dh[t] != t / 3 || g.tap_dw[t] != t % 3) return false;
  return true;
}

static int launch_wgrad_om(const TapGemm& g, const void* x, const void* dz, int O, float* partial, int splits,
                           cudaStream_t s) {
  CUtensorMap tmDz, tmX, tmDzS;
  const long long M = (long long)g.N * g.P * g.Q;
  PP_TRY(make_map_2d(&tmDz, dz, (uint64_t)M, (uint64_t)O, kWK));
  const bool paired = wgrad_om_paired(g, O);
  tmDzS = tmDz;
  if (paired) {
    TapGemm gz = g;   // dz viewed as an NHWC image [N, P, Q, O]
    gz.C = O; gz.H = g.P; gz.W = g.Q;
    PP_TRY(make_map_tiled4d(&tmDzS, dz, gz, g.Q, kWK / g.Q, 1));
  }
  int bw = 0, bh = 0, bn = 0;
  const bool x_tiled = prefer_tiled() && tiled_box_for(g, kWK, &bw, &bh, &bn);
  if (x_tiled) PP_TRY(make_map_tiled4d(&tmX, x, g, bw, bh, bn));
  else PP_TRY(make_map_im2col(&tmX, x, g, kWK));
  WgradOmDev p;
  p.x_tiled = x_tiled ? 1 : 0;
  p.M = (int)M; p.P = g.P; p.Q = g.Q; p.PQ = g.P * g.Q;
  p.base_h = g.base_h; p.base_w = g.base_w; p.step_h = g.step_h; p.step_w = g.step_w;
  p.C = g.C; p.O = O; p.ntaps = g.ntaps; p.Ktot = g.ntaps * g.C; p.nslabs = p.Ktot / 64;
  p.paired = paired ? 1 : 0;
  p.chunks_total = (int)((M + kWK - 1) / kWK);
  p.chunks_per_split = (p.chunks_total + splits - 1) / splits;
  for (int t = 0; t < g.ntaps; ++t) { p.tap_dh[t] = g.tap_dh[t]; p.tap_dw[t] = g.tap_dw[t]; }
  p.partial = partial;
  PP_SET_MAX_SMEM_ONCE((wgrad_om_kernel), kOmSmemBytes);
  dim3 grid(paired ? 2 : (p.nslabs + 3) / 4, splits);
  prof_begin(PROF_WGRAD, 2.0 * (double)M * O * g.ntaps * g.C, g.C, O, g.ntaps, s);
  wgrad_om_kernel<<<grid, kThreads, kOmSmemBytes, s>>>(tmX, tmDz, tmDzS, p);
  prof_end(PROF_WGRAD, s);
  PP_POST_LAUNCH();
  return PP_OK;
}

int wgrad_tcgen05(const TapGemm& g, const void* x, const void* dz, int O, float* partial, int splits,
                  cudaStream_t s) {
  PP_TRY(resolve_encoders());
  PP_REQUIRE(wgrad_tcgen05_supported(g, O), PP_EUNSUPPORTED, "tcgen05 wgrad needs C%%64==0 and O%%64==0 (C=%d O=%d)",
             g.C, O);
  if (g.tf32) {
    switch (wgrad_bn(O)) {
      case 256: return launch_wgrad<256, true>(g, x, dz, O, partial, splits, s);
      case 192: return launch_wgrad<192, true>(g, x, dz, O, partial, splits, s);
      case 128: return launch_wgrad<128, true>(g, x, dz, O, partial, splits, s);
      default: return launch_wgrad<64, true>(g, x, dz, O, partial, splits, s);
    }
  }
  if (wgrad_om_applies(g, O)) return launch_wgrad_om(g, x, dz, O, partial, splits, s);
  switch (wgrad_bn(O)) {
    case 256: return launch_wgrad<256>(g, x, dz, O, partial, splits, s);
    case 192: return launch_wgrad<192>(g, x, dz, O, partial, splits, s);
    case 128: return launch_wgrad<128>(g, x, dz, O, partial, splits, s);
    default: return launch_wgrad<64>(g, x, dz, O, partial, splits, s);
  }
}

int debug_last_timeout() {
  int v = 0;
  cudaMemcpyFromSymbol(&v, g_pp_timeout_code, sizeof(int));
  return v;
}

}  // namespace pp
