// Shared helpers for libdsb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/dsb200.h"

namespace dsb200 {

void set_error(const char* msg);
void count_launch(int n = 1);     // host-side tally of kernel launches issued by this library

#define DSB_REQUIRE(cond, msg) \
    do { if (!(cond)) { ::dsb200::set_error(msg); return -1; } } while (0)
#define DSB_LAUNCH_CHECK() \
    do { ::dsb200::count_launch(); cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) { ::dsb200::set_error(cudaGetErrorString(e__)); return (int)e__; } return 0; } while (0)

// Tuning knobs: read from the environment ONCE per process (first use), never per launch.  Knobs that switch parts of the
// work off or dump debug data (timing experiments: wrong results) exist only in -DDSB200_EXPERIMENTS builds.
int env_int_raw(const char* name, int dflt);
#define DSB_ENV_INT(name, dflt) \
    ([&]() -> int { static const int v__ = ::dsb200::env_int_raw(name, -2147483647); return v__ == -2147483647 ? (int)(dflt) : v__; }())
#ifdef DSB200_EXPERIMENTS
#define DSB_EXPERIMENT_INT(name, dflt) DSB_ENV_INT(name, dflt)
#else
#define DSB_EXPERIMENT_INT(name, dflt) (dflt)
#endif

constexpr int kNumSMs = 148;       // B200: 2 dies x 74 SMs

static inline int grid_for(long long work_items, int threads, int ctas_per_sm = 8) {
    long long blocks = (work_items + threads - 1) / threads;
    long long cap = (long long)kNumSMs * ctas_per_sm;
    if (blocks > cap) blocks = cap;          // grid-stride loops cover the rest
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float act_fwd(float z, int act) {
    switch (act) {
        case DSB200_ACT_RELU: return z > 0.f ? z : 0.f;
        case DSB200_ACT_ELU: return z > 0.f ? z : expm1f(z);
        case DSB200_ACT_TANH: return tanhf(z);
        case DSB200_ACT_SIGMOID: return 1.f / (1.f + expf(-z));
        case DSB200_ACT_LEAKY_RELU: return z > 0.f ? z : 0.2f * z;
        case DSB200_ACT_SOFTPLUS: return z > 20.f ? z : log1pf(expf(z));
        case DSB200_ACT_RELU6: return fminf(fmaxf(z, 0.f), 6.f);
        default: return z;
    }
}
// derivative of the activation w.r.t. its input z
__device__ __forceinline__ float act_grad(float z, int act) {
    switch (act) {
        case DSB200_ACT_RELU: return z > 0.f ? 1.f : 0.f;
        case DSB200_ACT_ELU: return z > 0.f ? 1.f : expf(z);
        case DSB200_ACT_TANH: { float t = tanhf(z); return 1.f - t * t; }
        case DSB200_ACT_SIGMOID: { float s = 1.f / (1.f + expf(-z)); return s * (1.f - s); }
        case DSB200_ACT_LEAKY_RELU: return z > 0.f ? 1.f : 0.2f;
        case DSB200_ACT_SOFTPLUS: return 1.f / (1.f + expf(-z));
        case DSB200_ACT_RELU6: return (z > 0.f && z < 6.f) ? 1.f : 0.f;
        default: return 1.f;
    }
}

// compile-time activation (ACT >= 0) or run-time dispatch (ACT < 0): the run-time switch costs ~30 instructions
// per element, which makes the elementwise kernels issue-bound
template <int ACT> __device__ __forceinline__ float act_fwd_t(float z, int act) {
    if (ACT == DSB200_ACT_RELU) return fmaxf(z, 0.f);
    if (ACT == DSB200_ACT_IDENTITY) return z;
    return act_fwd(z, act);
}
template <int ACT> __device__ __forceinline__ float act_grad_t(float z, int act) {
    if (ACT == DSB200_ACT_RELU) return z > 0.f ? 1.f : 0.f;
    if (ACT == DSB200_ACT_IDENTITY) return 1.f;
    return act_grad(z, act);
}

// Threads of a block are laid out [rows][G] (G = feature groups, a power of two): lanes whose index is equal
// modulo G hold partial sums of the same features.  Sum v over those lanes (G < 32); afterwards lanes 0..G-1
// of every warp hold the warp totals.  Cuts same-address shared-memory atomics by 32/G.
__device__ __forceinline__ float warp_sum_mod(float v, int G) {
    for (int o = 16; o >= G; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace dsb200
