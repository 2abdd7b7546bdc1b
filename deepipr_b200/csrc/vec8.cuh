// vec8.cuh — 8-channel (one 128-bit bf16 vector) load/store helpers shared by the HBM-bound passes
// (pointwise.cu, groupnorm.cu).  NHWC rows are O channels wide with O % 8 == 0.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace pp {

__device__ __forceinline__ void load8(const void* z, int z_f32, size_t vec_idx, float (&v)[8]) {
  if (z_f32) {
    const float4* p = reinterpret_cast<const float4*>(z) + vec_idx * 2;
    const float4 x0 = p[0], x1 = p[1];
    v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w;
    v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
  } else {
    const uint4 u = reinterpret_cast<const uint4*>(z)[vec_idx];
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
}
__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, size_t vec_idx, float (&v)[8]) {
  load8(p, 0, vec_idx, v);
}
__device__ __forceinline__ void store8_bf16(__nv_bfloat16* p, size_t vec_idx, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  reinterpret_cast<uint4*>(p)[vec_idx] = u;
}
// activation-typed store: fp32 (two 128-bit stores) when f32 != 0, else one bf16 vector
__device__ __forceinline__ void store8(void* p, int f32, size_t vec_idx, const float (&v)[8]) {
  if (f32) {
    float4* q = reinterpret_cast<float4*>(p) + vec_idx * 2;
    q[0] = make_float4(v[0], v[1], v[2], v[3]);
    q[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    store8_bf16(reinterpret_cast<__nv_bfloat16*>(p), vec_idx, v);
  }
}
__device__ __forceinline__ void load8_coef(const float* c, int ch, float (&v)[8]) {
  const float4 x0 = __ldg(reinterpret_cast<const float4*>(c + ch));
  const float4 x1 = __ldg(reinterpret_cast<const float4*>(c + ch + 4));
  v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w;
  v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
}

}  // namespace pp
