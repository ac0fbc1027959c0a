// Experiment: how does TMA lay out a 3-D box {16, 2, 8} of fp32 (inner 64 B) in shared memory under
// SWIZZLE_128B_ATOM_32B / SWIZZLE_128B / NONE?  Dumps the 8 x 128 B region.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -I deepsphere_tf1_b200/csrc -o /tmp/tma_box tools/experiments/tma_box_layout.cu -lcuda
#include <cstdio>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace dsb200::tc;

__global__ void k(const __grid_constant__ CUtensorMap tm, float* out, int nfloat)
{
    __shared__ __align__(1024) float buf[2048];
    __shared__ uint64_t bar;
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = -1.f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, 16 * 2 * 8 * 4);
        tma_load_3d(buf, &tm, &bar, 0, 0, 0);
    }
    mbar_wait(&bar, 0, 0);
    for (int i = threadIdx.x; i < nfloat; i += blockDim.x) out[i] = buf[i];
}

int main()
{
    const int Fin = 16, P = 4, R = 64;
    std::vector<float> h((size_t)P * R * Fin);
    for (int p = 0; p < P; ++p) for (int r = 0; r < R; ++r) for (int f = 0; f < Fin; ++f) h[((size_t)p * R + r) * Fin + f] = p * 10000 + r * 100 + f;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&o, 2048 * 4);
    void* fn; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = (EncodeTiledFn)fn;
    CUtensorMapSwizzle modes[3] = {CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B};
    for (int m = 0; m < 3; ++m) {
        CUtensorMap tm;
        cuuint64_t dims[3] = {Fin, P, R};
        cuuint64_t strides[2] = {(cuuint64_t)R * Fin * 4, (cuuint64_t)Fin * 4};
        cuuint32_t box[3] = {16, 2, 8}, es[3] = {1, 1, 1};
        CUresult rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, modes[m],
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("mode %d encode rc=%d\n", m, (int)rc);
        if (rc) continue;
        k<<<1, 128>>>(tm, o, 512);
        cudaError_t e = cudaDeviceSynchronize();
        printf("kernel: %s\n", cudaGetErrorString(e));
        std::vector<float> r(512);
        cudaMemcpy(r.data(), o, 512 * 4, cudaMemcpyDeviceToHost);
        for (int row = 0; row < 16; ++row) {
            printf("smem +%4d:", row * 128);
            for (int c = 0; c < 32; c += 4) printf(" %6.0f", r[row * 32 + c]);
            printf("\n");
        }
    }
    return 0;
}
