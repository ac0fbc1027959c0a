// Microbenchmark: TMA (cp.async.bulk.tensor.2d) throughput against the box shape.  One producer thread per CTA keeps S boxes
// in flight over a ring of mbarriers; a consumer warp only waits and releases.  148 CTAs (1 per SM), fp32 [rows x cols] tensor,
// box = [box_rows x box_cols], hardware swizzle matching the row width.  Prints GB/s per SM and aggregate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_rate tma_rate.cu -lcuda && ./tma_rate
#include "../../deepsphere_tf1_b200/csrc/tc_common.cuh"
#include <cstdio>
using namespace dsb200::tc;

static EncodeTiledFn enc() {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    return (EncodeTiledFn)p;
}

__global__ void __launch_bounds__(256) k(const __grid_constant__ CUtensorMap tm, int box_rows, int box_bytes, int col_boxes, long long rows, int S, long long* out, int P)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    __shared__ uint64_t fullA[64], emptyA[64];
    if (threadIdx.x == 0) { for (int s = 0; s < S * P; ++s) { mbar_init(fullA + s, 1); mbar_init(emptyA + s, 1); } fence_barrier_init(); }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pid = warp < P ? warp : warp - P;
    uint64_t* full = fullA + pid * S; uint64_t* empty = emptyA + pid * S;
    smem += (size_t)pid * S * box_bytes;
    const long long nboxes_total = (rows / box_rows) * col_boxes;
    const long long per = nboxes_total / gridDim.x / P;
    const long long b0 = ((long long)blockIdx.x * P + pid) * per;
    long long t0 = 0, t1 = 0;
    if (lane == 0 && warp < P) {
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        int s = 0; uint32_t ph = 0;
        for (long long i = 0; i < per; ++i) {
            const long long b = b0 + i;
            const int cb = (int)(b % col_boxes); const long long rb = b / col_boxes;
            mbar_wait(empty + s, ph ^ 1u, 1);
            mbar_expect_tx(full + s, box_bytes);
            tma_load_2d(smem + (size_t)s * box_bytes, &tm, full + s, cb * (box_bytes / box_rows / 4), (int)(rb * box_rows));
            if (++s == S) { s = 0; ph ^= 1u; }
        }
    } else if (lane == 0 && warp >= P && warp < 2 * P) {
        int s = 0; uint32_t ph = 0;
        for (long long i = 0; i < per; ++i) {
            mbar_wait(full + s, ph, 2);
            mbar_arrive(empty + s);
            if (++s == S) { s = 0; ph ^= 1u; }
        }
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        if (pid == 0) out[blockIdx.x] = t1;
    }
    if (threadIdx.x == 0) out[gridDim.x + blockIdx.x] = t0;
}

int main()
{
    const long long rows = 1 << 20;
    float* d; cudaMalloc(&d, rows * 128 * 4);
    cudaMemset(d, 0, rows * 128 * 4);
    long long* o; cudaMalloc(&o, 2 * 148 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Cfg { int cols, box_cols, box_rows; } cfgs[] = {
        {16, 16, 128}, {16, 16, 256}, {16, 16, 64}, {32, 32, 128}, {32, 32, 64}, {32, 32, 24}, {32, 32, 16},
        {64, 32, 128}, {64, 32, 64}, {128, 32, 128}, {128, 32, 32},
    };
    for (auto c : cfgs) {
        CUtensorMap tm;
        cuuint64_t dims[2] = {(cuuint64_t)c.cols, (cuuint64_t)rows};
        cuuint64_t strides[1] = {(cuuint64_t)c.cols * 4};
        cuuint32_t box[2] = {(cuuint32_t)c.box_cols, (cuuint32_t)c.box_rows};
        cuuint32_t es[2] = {1, 1};
        const CUtensorMapSwizzle sw = c.box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
        if (enc()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); continue; }
        const int box_bytes = c.box_cols * c.box_rows * 4;
        const int col_boxes = c.cols / c.box_cols;
        for (int P : {1, 2, 4}) for (int S : {4}) {
            if ((size_t)S * P * box_bytes > 190 * 1024) continue;
            k<<<148, 256, (size_t)S * P * box_bytes + 1024>>>(tm, c.box_rows, box_bytes, col_boxes, rows, S, o, P);
            long long h[296];
            if (cudaMemcpy(h, o, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) { printf("error\n"); return 1; }
            long long tmin = h[148], tmax = h[0];
            for (int i = 0; i < 148; ++i) { if (h[148 + i] < tmin) tmin = h[148 + i]; if (h[i] > tmax) tmax = h[i]; }
            const double bytes = (double)rows * c.cols * 4;
            const double ns = (double)(tmax - tmin);
            const double rows_per_sm = (double)rows * col_boxes / 148;
            printf("P=%d tensor [%lld x %3d] box %3d rows x %3d B, %d in flight (%3d KB/SM): %7.1f us  %6.0f GB/s  %5.2f ns per box row per SM\n",
                   P, rows, c.cols, c.box_rows, c.box_cols * 4, S * P, S * P * box_bytes / 1024, ns / 1e3, bytes / ns, ns / rows_per_sm);
        }
    }
    return 0;
}
