// K1: sparse-Laplacian products with the Chebyshev recurrence arithmetic fused in.
//   out[b] = alpha * L x[b] + beta * y[b] + gamma * z[b]   on [B, M, F] tensors (B samples).
// Replaces tf.sparse_tensor_dense_matmul + the `2*Lx1 - x0` elementwise ops of
// cgcnn.chebyshev5 (reference deepsphere/models.py:1056-1061) and, in the backward pass, the
// TF autodiff of the same lines (Clenshaw evaluation of sum_k T_k(L)^T G_k).
#include "common.cuh"
#include <atomic>

namespace dsb200 {

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static thread_local char g_err[256] = "";
void set_error(const char* msg) {
    int i = 0;
    for (; msg[i] && i < 255; ++i) g_err[i] = msg[i];
    g_err[i] = 0;
}

// One thread owns VEC consecutive features of one (sample, vertex) row; the threads of a row
// read the same (col, val) pairs (L1 broadcast) and gather 16-byte pieces of the neighbour rows
// of the same sample, so a warp touches whole 128-byte lines of every neighbour row when
// F >= 32.  Rows of consecutive vertices are spatially close on the sphere (HEALPix nested =
// Morton order), so the gathers mostly hit L1 / L2.  y / z / out are not __restrict__: out may
// alias y or z (in-place adjoint recurrence).
template <int VEC, bool CSR>
__global__ void __launch_bounds__(256)
spmm_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
            const float* __restrict__ val, int width, long long R, int M, int Fv,
            const float* __restrict__ x, const float* y, const float* z,
            float* out, float alpha, float beta, float gamma)
{
    const long long total = R * Fv;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        const long long r = g / Fv;
        const int fv = (int)(g - r * Fv);
        const int m = (int)(r % M);
        const long long base = r - m;                       // first row of this sample
        float acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
        int j0, j1;
        if (CSR) { j0 = __ldg(rowptr + m); j1 = __ldg(rowptr + m + 1); }
        else { j0 = m * width; j1 = j0 + width; }
        for (int j = j0; j < j1; ++j) {
            const int c = __ldg(col + j);
            if (!CSR && c < 0) break;                       // ELL padding sits at the end of a row
            const float w = __ldg(val + j);
            const float* xr = x + ((base + c) * Fv + fv) * VEC;
            if (VEC == 4) {
                const float4 xv = ldg4(xr);
                acc[0] = fmaf(w, xv.x, acc[0]); acc[1] = fmaf(w, xv.y, acc[1]);
                acc[2] = fmaf(w, xv.z, acc[2]); acc[3] = fmaf(w, xv.w, acc[3]);
            } else if (VEC == 2) {
                const float2 xv = __ldg(reinterpret_cast<const float2*>(xr));
                acc[0] = fmaf(w, xv.x, acc[0]); acc[1] = fmaf(w, xv.y, acc[1]);
            } else {
                acc[0] = fmaf(w, __ldg(xr), acc[0]);
            }
        }
        const long long o = g * VEC;
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] *= alpha;
        if (VEC == 4) {
            float4 rr = make_float4(acc[0], acc[1], acc[2], acc[3]);
            if (y) { const float4 t = *reinterpret_cast<const float4*>(y + o); rr.x = fmaf(beta, t.x, rr.x); rr.y = fmaf(beta, t.y, rr.y); rr.z = fmaf(beta, t.z, rr.z); rr.w = fmaf(beta, t.w, rr.w); }
            if (z) { const float4 t = *reinterpret_cast<const float4*>(z + o); rr.x = fmaf(gamma, t.x, rr.x); rr.y = fmaf(gamma, t.y, rr.y); rr.z = fmaf(gamma, t.z, rr.z); rr.w = fmaf(gamma, t.w, rr.w); }
            st4(out + o, rr);
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                float rr = acc[v];
                if (y) rr = fmaf(beta, y[o + v], rr);
                if (z) rr = fmaf(gamma, z[o + v], rr);
                out[o + v] = rr;
            }
        }
    }
}

static int spmm_launch(const int32_t* rowptr, const int32_t* col, const float* val, int width, int B, int M, int F,
                       const float* x, const float* y, const float* z, float* out,
                       float alpha, float beta, float gamma, cudaStream_t s)
{
    DSB_REQUIRE(B > 0 && M > 0 && F > 0, "spmm: empty matrix");
    DSB_REQUIRE(col && val && x && out, "spmm: null pointer");
    DSB_REQUIRE(x != out, "spmm: out must not alias x");
    if (beta == 0.f) y = nullptr;
    if (gamma == 0.f) z = nullptr;
    const uintptr_t align = (uintptr_t)x | (uintptr_t)out | (uintptr_t)y | (uintptr_t)z;
    const int vec = (F % 4 == 0 && (align & 15) == 0) ? 4 : (F % 2 == 0 && (align & 7) == 0) ? 2 : 1;
    const int Fv = F / vec;
    const long long R = (long long)B * M;
    const int grid = grid_for(R * Fv, 256, 8);
    if (!rowptr) DSB_REQUIRE(width > 0, "spmm_ell: width must be positive");
#define DSB_SPMM(V, C) spmm_kernel<V, C><<<grid, 256, 0, s>>>(rowptr, col, val, width, R, M, Fv, x, y, z, out, alpha, beta, gamma)
    if (rowptr) { if (vec == 4) DSB_SPMM(4, true); else if (vec == 2) DSB_SPMM(2, true); else DSB_SPMM(1, true); }
    else        { if (vec == 4) DSB_SPMM(4, false); else if (vec == 2) DSB_SPMM(2, false); else DSB_SPMM(1, false); }
#undef DSB_SPMM
    DSB_LAUNCH_CHECK();
}

}  // namespace dsb200

using namespace dsb200;

extern "C" int dsb200_abi_version(void) { return DSB200_ABI_VERSION; }
extern "C" const char* dsb200_last_error(void) { return g_err; }
extern "C" long long dsb200_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int dsb200_spmm_ell(const int32_t* col, const float* val, int width, int B, int M, int F,
                               const float* x, const float* y, const float* z, float* out,
                               float alpha, float beta, float gamma, void* stream)
{
    return spmm_launch(nullptr, col, val, width, B, M, F, x, y, z, out, alpha, beta, gamma, (cudaStream_t)stream);
}

extern "C" int dsb200_spmm_csr(const int32_t* rowptr, const int32_t* col, const float* val, int B, int M, int F,
                               const float* x, const float* y, const float* z, float* out,
                               float alpha, float beta, float gamma, void* stream)
{
    DSB_REQUIRE(rowptr, "spmm_csr: null rowptr");
    return spmm_launch(rowptr, col, val, 0, B, M, F, x, y, z, out, alpha, beta, gamma, (cudaStream_t)stream);
}

extern "C" int dsb200_cheb_basis(const int32_t* rowptr, const int32_t* col, const float* val, int width,
                                 int B, int M, int F, int K, const float* x0, float* basis, void* stream)
{
    DSB_REQUIRE(K >= 1, "cheb_basis: K must be >= 1");
    if (K == 1) return 0;
    DSB_REQUIRE(basis, "cheb_basis: null basis");
    const long long plane = (long long)B * M * F;
    cudaStream_t s = (cudaStream_t)stream;
    // X1 = L X0
    int rc = spmm_launch(rowptr, col, val, width, B, M, F, x0, nullptr, nullptr, basis, 1.f, 0.f, 0.f, s);
    // Xk = 2 L X(k-1) - X(k-2)
    for (int k = 2; k < K && rc == 0; ++k) {
        const float* xm1 = basis + (long long)(k - 2) * plane;
        const float* xm2 = (k == 2) ? x0 : basis + (long long)(k - 3) * plane;
        rc = spmm_launch(rowptr, col, val, width, B, M, F, xm1, xm2, nullptr, basis + (long long)(k - 1) * plane,
                         2.f, -1.f, 0.f, s);
    }
    return rc;
}

// Adjoint of X1 = L X0, Xk = 2 L X(k-1) - X(k-2):  with A_k = dLoss/dX_k (total),
//   A_(K-1) = G_(K-1);  A_k = G_k + 2 L^T A_(k+1) - A_(k+2)  (1 <= k <= K-2);
//   A_0 = G_0 + L^T A_1 - A_2.
extern "C" int dsb200_cheb_basis_bwd(const int32_t* rowptrT, const int32_t* colT, const float* valT, int width,
                                     int B, int M, int F, int K, float* G, void* stream)
{
    DSB_REQUIRE(K >= 1 && G, "cheb_basis_bwd: bad arguments");
    if (K == 1) return 0;
    const long long plane = (long long)B * M * F;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = 0;
    for (int k = K - 2; k >= 1 && rc == 0; --k) {
        float* gk = G + (long long)k * plane;
        const float* a1 = G + (long long)(k + 1) * plane;
        const float* a2 = (k + 2 <= K - 1) ? G + (long long)(k + 2) * plane : nullptr;
        rc = spmm_launch(rowptrT, colT, valT, width, B, M, F, a1, gk, a2, gk, 2.f, 1.f, a2 ? -1.f : 0.f, s);
    }
    if (rc) return rc;
    const float* a2 = (K > 2) ? G + 2 * plane : nullptr;
    return spmm_launch(rowptrT, colT, valT, width, B, M, F, G + plane, G, a2, G, 1.f, 1.f, a2 ? -1.f : 0.f, s);
}
