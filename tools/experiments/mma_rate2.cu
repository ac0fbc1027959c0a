// Microbenchmark 2: is the ~87 ns per tcgen05.mma of mma_rate.cu the tensor pipe or the issuing thread?
// Same instruction (cta_group::1, kind::tf32, M = 128, K = 8, operands K-major SW128 in shared memory), but the issue loop is
// fully unrolled with compile-time descriptor offsets (mma_rate.cu has a run-time modulo per MMA: ~40 SASS instructions),
// timed with clock64 on the issuing SM.  Variants: one accumulator / two alternating accumulators; N = 32..256.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_rate2 mma_rate2.cu && ./mma_rate2
#include "../../deepsphere_tf1_b200/csrc/tc_common.cuh"
#include <cstdio>
using namespace dsb200::tc;

template <int N, int ALT>
__global__ void __launch_bounds__(128) k(long long* out, int reps)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint64_t dA = make_smem_desc(smem_u32(smem), 16, 8u * 128u, 3);
        const uint64_t dB = make_smem_desc(smem_u32(smem) + 16384, 16, 8u * 128u, 3);
        const uint32_t idesc = make_idesc_tf32(128, N, 0, 0);
        uint32_t ph = 0;
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            for (int r = 0; r < reps; ++r) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    mma_tf32(tm + (ALT ? (uint32_t)((i & 1) * 256) : 0u), dA + 2u * (i & 3), dB + 2u * (i & 3), idesc, 1u);
            }
            const long long t1 = clock64();
            tc_commit(&bar);
            mbar_wait(&bar, ph, 0);
            ph ^= 1u;
            const long long t2 = clock64();
            out[rep * 2] = t1 - t0;
            out[rep * 2 + 1] = t2 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int N, int ALT> void run(long long* d)
{
    const int reps = 64;
    cudaFuncSetAttribute(k<N, ALT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    k<N, ALT><<<1, 128, 66 * 1024>>>(d, reps);
    long long h[6];
    cudaError_t e = cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
    const int n = reps * 16;
    printf("M=128 N=%3d alternate-accumulators=%d: %d MMAs issued in %lld cycles, complete in %lld cycles -> %.1f cycles per MMA issue, %.1f per MMA complete (floor M*N/256 = %d)\n",
           N, ALT, n, h[4], h[5], (double)h[4] / n, (double)h[5] / n, 128 * N / 256);
}

int main()
{
    long long* d;
    cudaMalloc(&d, 64);
    run<32, 0>(d); run<64, 0>(d); run<128, 0>(d); run<256, 0>(d);
    run<32, 1>(d); run<64, 1>(d); run<128, 1>(d); run<256, 1>(d);
    return 0;
}
