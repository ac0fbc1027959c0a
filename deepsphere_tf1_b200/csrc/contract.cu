// K2 (general fp32 path): the dense contraction of cgcnn.chebyshev5,
//   Y[r, fo] = sum_{k, fi} X_k[r, fi] * W[fi*K + k, fo]          (reference models.py:1062-1078)
// and the two products TF1's autodiff derives from it (models.py:831):
//   dX_k[r, fi] = sum_fo dY[r, fo] * W[fi*K + k, fo]
//   dW[fi*K + k, fo] = sum_r X_k[r, fi] * dY[r, fo]
// Rows are (sample, vertex) pairs of the [B, M, F] layout; the K Chebyshev orders are K
// separate [R, Fin] planes (never concatenated / transposed as the reference does at 1062-1064).
// This file: the dispatch of the three entry points and the small-reduction kernels of the input layer.  The tensor-core
// kernels live in contract_tc.cu / contract_tc2.cu / contract_mma.cu, the general fp32 FFMA GEMM (any shape) in contract_ffma.cu.
#include "common.cuh"
#include <stdlib.h>
#include "tc_common.cuh"
#include <atomic>

namespace dsb200 {


// Planar operand: logical column q of a [R, P*Fw] matrix lives in plane q / Fw (plane 0 at p0,
// planes >= 1 at pk + (plane-1)*R*Fw) at column q % Fw.
struct Planar {
    const float* p0;
    const float* pk;
    int Fw;
    long long plane_stride;
    __device__ __forceinline__ const float* at(int plane) const {
        return plane == 0 ? p0 : pk + (long long)(plane - 1) * plane_stride;
    }
};

// Small reduction depth (Q = K*Fin <= 16, e.g. the first layer Fin = 1): one thread produces 4
// output columns of one row straight from global memory; W sits in shared memory.  The kernel is
// bound by the Y store, so no tiling is needed.
__global__ void __launch_bounds__(256)
contract_small_fwd_kernel(Planar A, const float* __restrict__ W, float* __restrict__ Y,
                          long long R, int Fin, int K, int Fout, int Q, double* __restrict__ bn_sums)
{
    extern __shared__ float sm[];
    float* Ws = sm;                           // [Q][Fout], row q = k*Fin + fi
    float* red = sm + Q * Fout;               // [2][Fout]
    for (int i = threadIdx.x; i < Q * Fout; i += blockDim.x) {
        const int q = i / Fout, n = i - q * Fout;
        const int k = q / Fin, fi = q - k * Fin;
        Ws[i] = __ldg(W + ((long long)fi * K + k) * Fout + n);
    }
    if (bn_sums) for (int i = threadIdx.x; i < 2 * Fout; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    const int nv = Fout / 4;
    const int rows_per_iter = blockDim.x / nv;            // blockDim.x % nv == 0 by construction
    const int cv = threadIdx.x % nv, rl = threadIdx.x / nv;
    float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long r = (long long)blockIdx.x * rows_per_iter + rl; r < R; r += (long long)gridDim.x * rows_per_iter) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = 0; q < Q; ++q) {
            const int k = q / Fin, fi = q - k * Fin;
            const float a = __ldg(A.at(k) + r * Fin + fi);
            const float4 w = *reinterpret_cast<const float4*>(Ws + q * Fout + cv * 4);
            acc.x = fmaf(a, w.x, acc.x); acc.y = fmaf(a, w.y, acc.y); acc.z = fmaf(a, w.z, acc.z); acc.w = fmaf(a, w.w, acc.w);
        }
        st4(Y + r * Fout + cv * 4, acc);
        s[0] += acc.x; s[1] += acc.y; s[2] += acc.z; s[3] += acc.w;
        ss[0] = fmaf(acc.x, acc.x, ss[0]); ss[1] = fmaf(acc.y, acc.y, ss[1]);
        ss[2] = fmaf(acc.z, acc.z, ss[2]); ss[3] = fmaf(acc.w, acc.w, ss[3]);
    }
    if (bn_sums) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { atomicAdd(&red[cv * 4 + j], s[j]); atomicAdd(&red[Fout + cv * 4 + j], ss[j]); }
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * Fout; i += blockDim.x) atomicAdd(bn_sums + i, (double)red[i]);
    }
}

// dW for a small reduction depth: every thread keeps Q x 4 partial sums for its 4 columns while
// striding over rows; block-level shared-memory reduction, then one global atomic per entry.
template <int QMAX>
__global__ void __launch_bounds__(256)
contract_small_bwd_weight_kernel(Planar A, const float* __restrict__ dY, float* __restrict__ dW,
                                 long long R, int Fin, int K, int Fout, int Q)
{
    extern __shared__ float sm[];             // [Q][Fout]
    for (int i = threadIdx.x; i < Q * Fout; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    const int nv = Fout / 4;
    const int rows_per_iter = blockDim.x / nv;
    const int cv = threadIdx.x % nv, rl = threadIdx.x / nv;
    float acc[QMAX][4];
#pragma unroll
    for (int q = 0; q < QMAX; ++q) { acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f; }
    for (long long r = (long long)blockIdx.x * rows_per_iter + rl; r < R; r += (long long)gridDim.x * rows_per_iter) {
        const float4 g = ldg4(dY + r * Fout + cv * 4);
#pragma unroll
        for (int q = 0; q < QMAX; ++q) {
            if (q < Q) {
                const int k = q / Fin, fi = q - k * Fin;
                const float a = __ldg(A.at(k) + r * Fin + fi);
                acc[q][0] = fmaf(a, g.x, acc[q][0]); acc[q][1] = fmaf(a, g.y, acc[q][1]);
                acc[q][2] = fmaf(a, g.z, acc[q][2]); acc[q][3] = fmaf(a, g.w, acc[q][3]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < QMAX; ++q)
        if (q < Q) {
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(&sm[q * Fout + cv * 4 + j], acc[q][j]);
        }
    __syncthreads();
    for (int i = threadIdx.x; i < Q * Fout; i += blockDim.x) {
        const int q = i / Fout, n = i - q * Fout;
        const int k = q / Fin, fi = q - k * Fin;
        atomicAdd(dW + ((long long)fi * K + k) * Fout + n, sm[i]);
    }
}

// tensor-core (tcgen05) path on by default; dsb200_set_tensor_cores(0) forces the FP32 FFMA kernels
static std::atomic<int> g_use_tc{1};

static inline bool small_path(int Q, int Fout) { return Q <= 16 && Fout % 4 == 0 && Fout <= 256 && 256 % (Fout / 4) == 0; }

}  // namespace dsb200

using namespace dsb200;

extern "C" long long dsb200_contract_workspace_bytes(int Fin, int K, int Fout)
{
    return (long long)tc::contract_workspace_bytes(Fin, K, Fout);
}

namespace dsb200 {
int contract_fwd_mma(const float* x0, const float* xk, const float* W, float* Y, long long R, int Fin, int K, int Fout,
                     double* bn_sums, cudaStream_t s, float* Ypool = nullptr);   // contract_mma.cu
int contract_wgrad_mma(const float* x0, const float* xk, const float* dY, float* dW, long long R, int Fin, int K, int Fout,
                       cudaStream_t s);                                     // contract_mma.cu
int contract_fwd_ffma(const float* x0, const float* xk, const float* W, float* Y, long long R, int Fin, int K, int Fout,
                      double* bn_sums, cudaStream_t s);                     // contract_ffma.cu
int contract_bwd_data_ffma(const float* dY, const float* W, float* G, long long R, int Fin, int K, int Fout, cudaStream_t s);
int contract_bwd_weight_ffma(const float* x0, const float* xk, const float* dY, float* dW, long long R, int Fin, int K, int Fout,
                             cudaStream_t s);
static bool wgrad_mma_enabled() { return DSB_ENV_INT("DSB200_WGRAD_MMA", 1) != 0; }   // default on
static bool fwd_mma_enabled() { return DSB_ENV_INT("DSB200_FWD_MMA", 3) != 0; }   // default on
}  // namespace dsb200

extern "C" int dsb200_cheb_contract_fwd(const float* x0, const float* xk, const float* W, float* Y,
                                        long long R, int Fin, int K, int Fout, double* bn_sums, void* workspace,
                                        void* stream)
{
    DSB_REQUIRE(R > 0 && Fin > 0 && K > 0 && Fout > 0, "contract_fwd: bad shape");
    DSB_REQUIRE(x0 && W && Y && (K == 1 || xk), "contract_fwd: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    Planar A{x0, xk, Fin, R * (long long)Fin};
    const int Q = K * Fin;
    if (g_use_tc.load(std::memory_order_relaxed) && fwd_mma_enabled()) {
        const int rc = contract_fwd_mma(x0, xk, W, Y, R, Fin, K, Fout, bn_sums, s);   // narrow layers (Fin 16 -> Fout 32)
        if (rc != 1) return rc;
    }
    if (g_use_tc.load(std::memory_order_relaxed) && Q > 16) {
        int rc = tc::contract_fwd_tc_cols(x0, xk, W, Y, R, Fin, K, Fout, bn_sums, workspace, s);
        if (rc == 1) rc = tc::contract_fwd_tc(x0, xk, W, Y, R, Fin, K, Fout, bn_sums, workspace, s);
        if (rc != 1) return rc;                            // 1 = shape not supported on the tensor-core path
    }
    if (small_path(Q, Fout) && (((uintptr_t)Y & 15) == 0)) {
        const int rows_per_iter = 256 / (Fout / 4);
        const int grid = grid_for((R + rows_per_iter - 1) / rows_per_iter * 256, 256, 8);
        const size_t smem = (size_t)(Q * Fout + 2 * Fout) * sizeof(float);
        contract_small_fwd_kernel<<<grid, 256, smem, s>>>(A, W, Y, R, Fin, K, Fout, Q, bn_sums);
    } else {
        return contract_fwd_ffma(x0, xk, W, Y, R, Fin, K, Fout, bn_sums, s);
    }
    DSB_LAUNCH_CHECK();
}

// Forward contraction with the HEALPix max-pool over 4 consecutive vertices fused into the epilogue: Ypool [R/4, Fout] =
// max over every 4 consecutive rows of Y; Y itself is written only when the caller needs it (training: the batch-norm backward
// reads it) and may be NULL (inference).  Tensor-core kernels only: 1 -> the caller uses the unfused operators.
static int contract_fwd_pool4(const float* x0, const float* xk, const float* W, float* Y, float* Ypool, long long R, int Fin, int K,
                              int Fout, double* bn_sums, void* workspace, cudaStream_t s)
{
    if (!g_use_tc.load(std::memory_order_relaxed) || (R & 3)) return 1;
    if (fwd_mma_enabled()) {
        const int rc = contract_fwd_mma(x0, xk, W, Y, R, Fin, K, Fout, bn_sums, s, Ypool);
        if (rc != 1) return rc;
    }
    if (K * Fin <= 16 || !workspace) return 1;
    int rc = tc::contract_fwd_tc_cols(x0, xk, W, Y, R, Fin, K, Fout, bn_sums, workspace, s, Ypool);
    if (rc == 1) rc = tc::contract_fwd_tc(x0, xk, W, Y, R, Fin, K, Fout, bn_sums, workspace, s, Ypool);
    return rc;
}

extern "C" int dsb200_cheb_contract_fwd_pool4(const float* x0, const float* xk, const float* W, float* Y, float* Ypool,
                                              long long R, int Fin, int K, int Fout, double* bn_sums, void* workspace, void* stream)
{
    DSB_REQUIRE(R > 0 && Fin > 0 && K > 0 && Fout > 0, "contract_fwd_pool4: bad shape");
    DSB_REQUIRE(x0 && W && Ypool && (K == 1 || xk), "contract_fwd_pool4: null pointer");
    const int rc = contract_fwd_pool4(x0, xk, W, Y, Ypool, R, Fin, K, Fout, bn_sums, workspace, (cudaStream_t)stream);
    DSB_REQUIRE(rc != 1, "contract_fwd_pool4: no tensor-core kernel takes this shape (see dsb200_contract_pool4_supported)");
    return rc;
}

// 1 when dsb200_cheb_contract_fwd_pool4 takes the shape (the conditions of the three tensor-core forward kernels)
extern "C" int dsb200_contract_pool4_supported(long long R, int Fin, int K, int Fout)
{
    if (!g_use_tc.load(std::memory_order_relaxed) || (R & 3) || R < 256 || R > 0x7fffff00LL) return 0;
    if (fwd_mma_enabled() && Fin == 16 && Fout == 32 && K >= 1 && K <= 8) return 1;
    if (K * Fin <= 16) return 0;
    if ((Fin & 7) || Fout > 256) return 0;
    return tc::contract_fwd_tc_query(R, Fin, K, Fout);
}

extern "C" int dsb200_cheb_contract_bwd_data(const float* dY, const float* W, float* G,
                                             long long R, int Fin, int K, int Fout, void* workspace, void* stream)
{
    DSB_REQUIRE(R > 0 && Fin > 0 && K > 0 && Fout > 0, "contract_bwd_data: bad shape");
    DSB_REQUIRE(dY && W && G, "contract_bwd_data: null pointer");
    if (g_use_tc.load(std::memory_order_relaxed)) {
        const int rc = tc::contract_bwd_data_tc(dY, W, G, R, Fin, K, Fout, workspace, (cudaStream_t)stream);
        if (rc != 1) return rc;
    }
    return contract_bwd_data_ffma(dY, W, G, R, Fin, K, Fout, (cudaStream_t)stream);
}

extern "C" int dsb200_cheb_contract_bwd_weight(const float* x0, const float* xk, const float* dY, float* dW,
                                               long long R, int Fin, int K, int Fout, void* stream)
{
    DSB_REQUIRE(R > 0 && Fin > 0 && K > 0 && Fout > 0, "contract_bwd_weight: bad shape");
    DSB_REQUIRE(x0 && dY && dW && (K == 1 || xk), "contract_bwd_weight: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    Planar A{x0, xk, Fin, R * (long long)Fin};
    const int Q = K * Fin;
    if (g_use_tc.load(std::memory_order_relaxed) && wgrad_mma_enabled()) {
        const int rc = contract_wgrad_mma(x0, xk, dY, dW, R, Fin, K, Fout, s);       // narrow layers (Fin 16 -> Fout 32)
        if (rc != 1) return rc;
    }
    if (g_use_tc.load(std::memory_order_relaxed) && Q > 16) {
        const int rc = tc::contract_bwd_weight_tc(x0, xk, dY, dW, R, Fin, K, Fout, s);
        if (rc != 1) return rc;
    }
    if (small_path(Q, Fout) && (((uintptr_t)dY & 15) == 0)) {
        const int rows_per_iter = 256 / (Fout / 4);
        const int grid = grid_for((R + rows_per_iter - 1) / rows_per_iter * 256, 256, 4);
        const size_t smem = (size_t)Q * Fout * sizeof(float);
        if (Q <= 8) contract_small_bwd_weight_kernel<8><<<grid, 256, smem, s>>>(A, dY, dW, R, Fin, K, Fout, Q);
        else contract_small_bwd_weight_kernel<16><<<grid, 256, smem, s>>>(A, dY, dW, R, Fin, K, Fout, Q);
    } else {
        return contract_bwd_weight_ffma(x0, xk, dY, dW, R, Fin, K, Fout, s);
    }
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_set_tensor_cores(int on)
{
    return g_use_tc.exchange(on ? 1 : 0);
}
