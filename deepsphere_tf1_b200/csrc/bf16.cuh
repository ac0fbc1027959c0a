// BF16 storage helpers of the reduced-precision mode (north star: "1e-2 in BF16"): activations are STORED as bfloat16
// ([B, M, F] tensors, Chebyshev planes, gradients of activations); every kernel unpacks to fp32, computes and accumulates
// in fp32 and rounds once (to nearest even) when it stores.  Weights, batch-norm statistics and weight gradients stay fp32.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace dsb200 {

typedef __nv_bfloat16 bf16;

// a 32-bit word holds two bf16: element 0 in the low half
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);          // .x = a -> low half
    return *reinterpret_cast<const uint32_t*>(&t);
}
__device__ __forceinline__ float4 unpack_bf16x4(uint2 w) { return make_float4(bf16_lo(w.x), bf16_hi(w.x), bf16_lo(w.y), bf16_hi(w.y)); }
__device__ __forceinline__ uint2 pack_bf16x4(float4 v) { return make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w)); }

// overloads of the fp32 helpers in common.cuh: 4 consecutive elements (8 bytes, 8-byte aligned), one element
__device__ __forceinline__ float4 ldg4(const bf16* p) { return unpack_bf16x4(__ldg(reinterpret_cast<const uint2*>(p))); }
__device__ __forceinline__ void st4(bf16* p, float4 v) { *reinterpret_cast<uint2*>(p) = pack_bf16x4(v); }
__device__ __forceinline__ float ldg1(const bf16* p) { return bf16_lo((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p))); }
__device__ __forceinline__ float ldg1(const float* p) { return __ldg(p); }
__device__ __forceinline__ void st1(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }

}  // namespace dsb200
