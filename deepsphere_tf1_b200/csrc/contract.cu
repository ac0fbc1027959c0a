// K2 (general fp32 path): the dense contraction of cgcnn.chebyshev5,
//   Y[r, fo] = sum_{k, fi} X_k[r, fi] * W[fi*K + k, fo]          (reference models.py:1062-1078)
// and the two products TF1's autodiff derives from it (models.py:831):
//   dX_k[r, fi] = sum_fo dY[r, fo] * W[fi*K + k, fo]
//   dW[fi*K + k, fo] = sum_r X_k[r, fi] * dY[r, fo]
// Rows are (sample, vertex) pairs of the [B, M, F] layout; the K Chebyshev orders are K
// separate [R, Fin] planes (never concatenated / transposed as the reference does at 1062-1064).
// FP32 FFMA register-tiled kernels (exact-precision path); the tcgen05 path lives in
// contract_tc.cu.
#include "common.cuh"
#include <stdlib.h>
#include "tc_common.cuh"
#include <atomic>

namespace dsb200 {

constexpr int BM = 128, BN = 64, BK = 16;      // CTA tile; 256 threads, 8x4 outputs per thread
constexpr int PADM = BM + 4, PADN = BN + 4;

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

// C[r, n] = sum_q A[r, q] * Bm[q, n]
//  FWD : A planar (X_k, Fw = Fin, Q = K*Fin, q = k*Fin + fi); Bm[q, n] = W[(fi*K + k)*Fout + n];
//        C = Y [R, Fout]; optional batch-norm sums in the epilogue.
//  BWD : A = dY [R, Fout] (Q = Fout); n = k*Fin + fi; Bm[q, n] = W[(fi*K + k)*Fout + q];
//        C planar (G_k [R, Fin]).
template <bool BWD>
__global__ void __launch_bounds__(256)
contract_kernel(Planar A, const float* __restrict__ W, float* __restrict__ Cout,
                long long R, int Fin, int K, int Fout, int Q, int N, double* __restrict__ bn_sums)
{
    __shared__ __align__(16) float As[BK][PADM];
    __shared__ __align__(16) float Bs[BK][PADN];
    __shared__ float red[2][BN];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;              // 16 x 16 threads
    const long long r0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int Aw = A.Fw;                                   // width of one A plane

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const bool a_vec = (Aw % 4 == 0);
    for (int q0 = 0; q0 < Q; q0 += BK) {
        // ---- A tile: BM rows x BK columns -> As[kk][row]
        if (a_vec) {
            // 512 float4 per tile: thread loads 2 (rows i and i+64)
            const int kq = (tid & 3) * 4;
            const int q = q0 + kq;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = (tid >> 2) + h * 64;
                const long long r = r0 + i;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < R && q < Q) {
                    const int plane = q / Aw, f = q - plane * Aw;
                    v = ldg4(A.at(plane) + r * Aw + f);
                }
                As[kq + 0][i] = v.x; As[kq + 1][i] = v.y; As[kq + 2][i] = v.z; As[kq + 3][i] = v.w;
            }
        } else {
            const int kk = tid & 15;
            const int q = q0 + kk;
            int plane = 0, f = 0;
            if (q < Q) { plane = q / Aw; f = q - plane * Aw; }
            const float* ap = A.at(plane);
#pragma unroll
            for (int h = 0; h < 8; ++h) {
                const int i = (tid >> 4) + h * 16;
                const long long r = r0 + i;
                As[kk][i] = (r < R && q < Q) ? __ldg(ap + r * Aw + f) : 0.f;
            }
        }
        // ---- B tile: BK x BN -> Bs[kk][n]
        {
            const int kk = tid >> 4;
            const int nn = (tid & 15) * 4;
            const int q = q0 + kk;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (q < Q) {
                if (!BWD) {
                    const int k = q / Fin, fi = q - k * Fin;
                    const float* wp = W + ((long long)fi * K + k) * Fout + n0 + nn;
                    if ((Fout % 4 == 0) && n0 + nn + 3 < N) {
                        const float4 t = ldg4(wp);
                        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (n0 + nn + j < N) v[j] = __ldg(wp + j);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int n = n0 + nn + j;
                        if (n < N) {
                            const int k = n / Fin, fi = n - k * Fin;
                            v[j] = __ldg(W + ((long long)fi * K + k) * Fout + q);
                        }
                    }
                }
            }
            *reinterpret_cast<float4*>(&Bs[kk][nn]) = make_float4(v[0], v[1], v[2], v[3]);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue
    const int nbase = n0 + tx * 4;
    if (!BWD) {
        const bool vec = (Fout % 4 == 0) && (nbase + 3 < N);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long r = r0 + ty * 8 + i;
            if (r >= R) continue;
            float* cp = Cout + r * Fout + nbase;
            if (vec) st4(cp, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (nbase + j < N) cp[j] = acc[i][j];
            }
        }
        if (bn_sums) {                                     // uniform branch
            if (tid < BN) { red[0][tid] = 0.f; red[1][tid] = 0.f; }
            __syncthreads();
            float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (r0 + ty * 8 + i < R) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { s[j] += acc[i][j]; ss[j] = fmaf(acc[i][j], acc[i][j], ss[j]); }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) { atomicAdd(&red[0][tx * 4 + j], s[j]); atomicAdd(&red[1][tx * 4 + j], ss[j]); }
            __syncthreads();
            if (tid < BN && n0 + tid < N) {
                atomicAdd(bn_sums + n0 + tid, (double)red[0][tid]);
                atomicAdd(bn_sums + N + n0 + tid, (double)red[1][tid]);
            }
        }
    } else {
        // n = k*Fin + fi  ->  plane k, column fi of G
        const bool vec = (Fin % 4 == 0) && (nbase + 3 < N);
        const long long plane_stride = R * (long long)Fin;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long r = r0 + ty * 8 + i;
            if (r >= R) continue;
            if (vec) {
                const int k = nbase / Fin, fi = nbase - k * Fin;
                st4(Cout + k * plane_stride + r * Fin + fi, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = nbase + j;
                    if (n < N) { const int k = n / Fin, fi = n - k * Fin; Cout[k * plane_stride + r * Fin + fi] = acc[i][j]; }
                }
            }
        }
    }
}

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

// dW tile kernel: rows of the output are q = k*Fin + fi, columns fo; the reduction runs over the
// (vertex, sample) rows r, split across blockIdx.z; partial tiles are merged with atomics.
constexpr int WM = 64, WN = 64, WK = 16;
__global__ void __launch_bounds__(256)
contract_bwd_weight_kernel(Planar A, const float* __restrict__ dY, float* __restrict__ dW,
                           long long R, int Fin, int K, int Fout, int Q, long long rows_per_split)
{
    __shared__ __align__(16) float As[WK][WM + 4];
    __shared__ __align__(16) float Bs[WK][WN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int q0 = blockIdx.x * WM, n0 = blockIdx.y * WN;
    const long long rbeg = (long long)blockIdx.z * rows_per_split;
    long long rend = rbeg + rows_per_split;
    if (rend > R) rend = R;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int li = tid & 63, lk = tid >> 6;               // 64 columns x 4 rows per pass
    const int qa = q0 + li;
    int plane = 0, fa = 0;
    if (qa < Q) { plane = qa / Fin; fa = qa - plane * Fin; }
    const float* ap = A.at(plane);
    const int nb = n0 + li;

    for (long long rr = rbeg; rr < rend; rr += WK) {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const int kk = lk + h * 4;
            const long long r = rr + kk;
            const bool ok = r < rend;
            As[kk][li] = (ok && qa < Q) ? __ldg(ap + r * Fin + fa) : 0.f;
            Bs[kk][li] = (ok && nb < Fout) ? __ldg(dY + r * Fout + nb) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < WK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = q0 + ty * 4 + i;
        if (q >= Q) continue;
        const int k = q / Fin, fi = q - k * Fin;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < Fout) atomicAdd(dW + ((long long)fi * K + k) * Fout + n, acc[i][j]);
        }
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
        dim3 grid((unsigned)((R + BM - 1) / BM), (unsigned)((Fout + BN - 1) / BN));
        contract_kernel<false><<<grid, 256, 0, s>>>(A, W, Y, R, Fin, K, Fout, Q, Fout, bn_sums);
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
    Planar A{dY, nullptr, Fout, 0};
    const int N = K * Fin;
    dim3 grid((unsigned)((R + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
    contract_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(A, W, G, R, Fin, K, Fout, Fout, N, nullptr);
    DSB_LAUNCH_CHECK();
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
        const int tq = (Q + WM - 1) / WM, tn = (Fout + WN - 1) / WN;
        long long splits = (4LL * kNumSMs + (long long)tq * tn - 1) / ((long long)tq * tn);
        const long long max_splits = (R + 4 * WK - 1) / (4 * WK);
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
        if (splits > 65535) splits = 65535;
        long long rows_per_split = (R + splits - 1) / splits;
        rows_per_split = (rows_per_split + WK - 1) / WK * WK;
        splits = (R + rows_per_split - 1) / rows_per_split;
        dim3 grid((unsigned)tq, (unsigned)tn, (unsigned)splits);
        contract_bwd_weight_kernel<<<grid, 256, 0, s>>>(A, dY, dW, R, Fin, K, Fout, Q, rows_per_split);
    }
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_set_tensor_cores(int on)
{
    return g_use_tc.exchange(on ? 1 : 0);
}
