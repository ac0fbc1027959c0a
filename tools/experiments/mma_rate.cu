// Microbenchmark: issue rate / latency of tcgen05.mma kind::tf32 (M = 128, K = 8 per instruction) for several N,
// operands K-major in shared memory (contents irrelevant), accumulating into one TMEM tile.  One CTA, one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_rate mma_rate.cu && ./mma_rate
#include "../../deepsphere_tf1_b200/csrc/tc_common.cuh"
#include <cstdio>
using namespace dsb200::tc;

__global__ void __launch_bounds__(128) k(long long* out, int N, int nmma, int kc, int swz, int distinct, int Mt)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const int rowbytes = kc * 4;
        const uint64_t dA = make_smem_desc(smem_u32(smem), 16, 8u * rowbytes, swz);
        const uint64_t dB = make_smem_desc(smem_u32(smem) + 16384, 16, 8u * rowbytes, swz);
        const uint32_t idesc = make_idesc_tf32(Mt, N, 0, 0);
        uint32_t ph = 0;
        for (int rep = 0; rep < 3; ++rep) {
            long long t0, t1, t2;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
            for (int i = 0; i < nmma; ++i)
                mma_tf32(tm + (distinct ? (uint32_t)((i & 1) * 256) : 0u), dA + 2u * (i % (kc / 8)), dB + 2u * (i % (kc / 8)), idesc, i > 0);
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
            tc_commit(&bar);
            mbar_wait(&bar, ph, 0);
            ph ^= 1u;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t2));
            out[rep * 2] = t1 - t0;
            out[rep * 2 + 1] = t2 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main()
{
    long long* d;
    cudaMalloc(&d, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int Ns[] = {32, 64, 128, 160, 256};
    for (int Mt : {128, 64})
    for (int kc : {32})
        for (int distinct : {0})
            for (int N : Ns) {
                const int nmma = 256;
                k<<<1, 128, 50 * 1024>>>(d, N, nmma, kc, kc == 32 ? 3 : 2, distinct, Mt);
                long long h[6];
                cudaError_t e = cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                printf("M=%d kc=%d alternate-accumulators=%d N=%3d: %d MMAs issued in %lld ns, complete in %lld ns -> %.1f ns per MMA (%.0f GFLOP/s per SM)\n",
                       Mt, kc, distinct, N, nmma, h[4], h[5], (double)h[5] / nmma, 2.0 * Mt * N * 8 * nmma / (double)h[5]);
            }
    return 0;
}
