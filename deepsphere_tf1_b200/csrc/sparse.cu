// K1: sparse-Laplacian products with the Chebyshev recurrence arithmetic fused in.
//   out[b] = alpha * L x[b] + beta * y[b] + gamma * z[b]   on [B, M, F] tensors (B samples).
// Replaces tf.sparse_tensor_dense_matmul + the `2*Lx1 - x0` elementwise ops of
// cgcnn.chebyshev5 (reference deepsphere/models.py:1056-1061) and, in the backward pass, the
// TF autodiff of the same lines (Clenshaw evaluation of sum_k T_k(L)^T G_k).
#include "common.cuh"
#include <atomic>
#include <stdlib.h>

namespace dsb200 {

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static thread_local char g_err[256] = "";
int env_int_raw(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

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
            if (!CSR && c < 0) break;                       // tolerate (col = -1) padding as well
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

// Main kernel (F % 4 == 0).  A CTA works on a compact TILE: TR consecutive vertex rows (a square patch of
// the sphere in HEALPix nested = Morton order) x a slab of at most 64 16-byte chunks (1 KB) of the feature
// row, so that the 9 gathers per row mostly hit the rows its neighbours in the same tile just fetched
// (L1-resident: about (sqrt(TR)+2)^2 distinct rows for 9*TR gathers).  Inside the tile a group of
// LG = 2^lg2 lanes owns one row and sweeps its chunks, PASSES chunks per lane: a warp gathers whole
// 128-byte lines as soon as the slab is >= 32 floats wide, the (col, val) pair of a neighbour is loaded
// once per lane and reused for PASSES gathers, and PASSES x (unrolled j) independent 16-byte loads are in
// flight per lane.  ELL padding slots hold (col = row, val = 0): uniform trip count, no data-dependent
// exit.  Tiles = (sample, row tile, column slab), distributed grid-stride over the CTAs.
// W > 0: ELL rows of exactly W slots, fully unrolled: all W (col, val) pairs are fetched first, then all
// W x PASSES gathers are issued back to back (one memory latency per row instead of one per neighbour).
template <int PASSES, int W, bool CSR>
__global__ void __launch_bounds__(256)
spmm_rows_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ val,
                 int width, int B, int M, int chunks, int lg2, int TR, int CB,
                 const float4* __restrict__ x, const float4* y, const float4* z, float4* out,
                 float alpha, float beta, float gamma)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int LG = 1 << lg2, rpw = 32 >> lg2;                    // lanes per row, rows per warp
    const int sub = lane >> lg2, lg = lane & (LG - 1);
    const int rtiles = (M + TR - 1) / TR, ctiles = (chunks + CB - 1) / CB;
    const long long tiles = (long long)B * rtiles * ctiles;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int ct = (int)(t % ctiles);
        const long long t2 = t / ctiles;
        const int rt = (int)(t2 % rtiles), b = (int)(t2 / rtiles);
        const long long sbase = (long long)b * M * chunks;        // first chunk of this sample
        const float4* xb = x + sbase + ct * CB + lg;
        const int mend = min(M, (rt + 1) * TR);
        bool on[PASSES];
#pragma unroll
        for (int p = 0; p < PASSES; ++p) on[p] = (ct * CB + lg + p * LG < chunks) && (lg + p * LG < CB);
        for (int m = rt * TR + warp * rpw + sub; m < mend; m += 8 * rpw) {
            float4 acc[PASSES];
#pragma unroll
            for (int p = 0; p < PASSES; ++p) acc[p] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (W > 0) {
                int c[W > 0 ? W : 1];
                float wv[W > 0 ? W : 1];
#pragma unroll
                for (int j = 0; j < W; ++j) { c[j] = __ldg(col + (long long)m * W + j); wv[j] = __ldg(val + (long long)m * W + j); }
                float4 v[W > 0 ? W : 1][PASSES];
#pragma unroll
                for (int j = 0; j < W; ++j) {
#pragma unroll
                    for (int p = 0; p < PASSES; ++p)
                        v[j][p] = on[p] ? __ldg(xb + (long long)c[j] * chunks + p * LG) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int j = 0; j < W; ++j) {
#pragma unroll
                    for (int p = 0; p < PASSES; ++p) {
                        acc[p].x = fmaf(wv[j], v[j][p].x, acc[p].x); acc[p].y = fmaf(wv[j], v[j][p].y, acc[p].y);
                        acc[p].z = fmaf(wv[j], v[j][p].z, acc[p].z); acc[p].w = fmaf(wv[j], v[j][p].w, acc[p].w);
                    }
                }
            } else {
                int j0, j1;
                if (CSR) { j0 = __ldg(rowptr + m); j1 = __ldg(rowptr + m + 1); }
                else { j0 = m * width; j1 = j0 + width; }
#pragma unroll 4
                for (int j = j0; j < j1; ++j) {
                    const int c = __ldg(col + j);
                    const float wv = __ldg(val + j);
                    const float4* xr = xb + (long long)c * chunks;
#pragma unroll
                    for (int p = 0; p < PASSES; ++p) {
                        if (on[p]) {
                            const float4 v = __ldg(xr + p * LG);
                            acc[p].x = fmaf(wv, v.x, acc[p].x); acc[p].y = fmaf(wv, v.y, acc[p].y);
                            acc[p].z = fmaf(wv, v.z, acc[p].z); acc[p].w = fmaf(wv, v.w, acc[p].w);
                        }
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < PASSES; ++p) {
                if (on[p]) {
                    const long long o = sbase + (long long)m * chunks + ct * CB + lg + p * LG;
                    float4 r = make_float4(acc[p].x * alpha, acc[p].y * alpha, acc[p].z * alpha, acc[p].w * alpha);
                    if (y) { const float4 t4 = y[o]; r.x = fmaf(beta, t4.x, r.x); r.y = fmaf(beta, t4.y, r.y); r.z = fmaf(beta, t4.z, r.z); r.w = fmaf(beta, t4.w, r.w); }
                    if (z) { const float4 t4 = z[o]; r.x = fmaf(gamma, t4.x, r.x); r.y = fmaf(gamma, t4.y, r.y); r.z = fmaf(gamma, t4.z, r.z); r.w = fmaf(gamma, t4.w, r.w); }
                    out[o] = r;
                }
            }
        }
    }
}

template <int PASSES>
static void launch_rows(const int32_t* rowptr, const int32_t* col, const float* val, int width, int B, int M, int chunks,
                        int lg2, int TR, int CB, const float* x, const float* y, const float* z, float* out,
                        float alpha, float beta, float gamma, cudaStream_t s)
{
    const long long tiles = (long long)B * ((M + TR - 1) / TR) * ((chunks + CB - 1) / CB);
    const long long cap = (long long)kNumSMs * 8;
    const int grid = (int)(tiles < cap ? tiles : cap);
#define DSB_ROWS(WW, CC) spmm_rows_kernel<PASSES, WW, CC><<<grid, 256, 0, s>>>(rowptr, col, val, width, B, M, chunks, lg2, TR, CB, \
            (const float4*)x, (const float4*)y, (const float4*)z, (float4*)out, alpha, beta, gamma)
    if (rowptr) DSB_ROWS(0, true);
    else if (width == 9) DSB_ROWS(9, false);
    else if (width == 7) DSB_ROWS(7, false);
    else DSB_ROWS(0, false);
#undef DSB_ROWS
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
    if (!rowptr) DSB_REQUIRE(width > 0, "spmm_ell: width must be positive");
    if (vec == 4) {
        const int chunks = F / 4;
        int CB = chunks < 64 ? chunks : 64;                        // column slab: at most 1 KB of a row
        { const int v = DSB_ENV_INT("DSB200_SPMM_CB", 0); if (v > 0 && v < CB) CB = v; }
        int lg2 = 0;
        while ((1 << lg2) < CB && lg2 < 5) ++lg2;
        const int passes = (CB + (1 << lg2) - 1) >> lg2;           // 1 or 2
        // rows per tile: ~64 KB of x per tile (it has to live in L1 next to its halo), 32..512 rows
        int TR = 32768 / (CB * 16);
        TR = TR < 32 ? 32 : TR > 128 ? 128 : TR;                   // 128-row tiles measured best on B200 (tools/sweep_spmm.py)
        { const int v = DSB_ENV_INT("DSB200_SPMM_TR", 0); if (v > 0) TR = v; }
        // keep enough tiles to fill the machine a few times over
        while (TR > 16 && (long long)B * ((M + TR - 1) / TR) * ((chunks + CB - 1) / CB) < 8LL * kNumSMs) TR >>= 1;
        if (passes <= 1) launch_rows<1>(rowptr, col, val, width, B, M, chunks, lg2, TR, CB, x, y, z, out, alpha, beta, gamma, s);
        else launch_rows<2>(rowptr, col, val, width, B, M, chunks, lg2, TR, CB, x, y, z, out, alpha, beta, gamma, s);
        DSB_LAUNCH_CHECK();
    }
    const int Fv = F / vec;
    const long long R = (long long)B * M;
    const int grid = grid_for(R * Fv, 256, 8);
#define DSB_SPMM(V, C) spmm_kernel<V, C><<<grid, 256, 0, s>>>(rowptr, col, val, width, R, M, Fv, x, y, z, out, alpha, beta, gamma)
    if (rowptr) { if (vec == 4) DSB_SPMM(4, true); else if (vec == 2) DSB_SPMM(2, true); else DSB_SPMM(1, true); }
    else        { if (vec == 4) DSB_SPMM(4, false); else if (vec == 2) DSB_SPMM(2, false); else DSB_SPMM(1, false); }
#undef DSB_SPMM
    DSB_LAUNCH_CHECK();
}


// ------------------------------------------------------------------------------------------------------------------
// Packed ELL: rec[j][m] = (col, val bits) of slot j of row m as ONE 8-byte record, slot-major ("transposed").  The
// rows a warp works on are consecutive, so slot j of all of them is one contiguous 8*rpw-byte piece: ONE L1 wavefront
// per slot and warp instead of the ~3 + ~3 of the row-major col[] / val[] arrays (rows 36 bytes apart).  At the 64- and
// 128-byte feature rows of the cosmo layers the (col, val) loads were half of all L1 wavefronts of the gather kernel
// (profiles/r01_spmm_rows_layer2_full.txt: L1/TEX 74 % busy, DRAM 38 %).
__global__ void ell_pack_kernel(const int32_t* __restrict__ col, const float* __restrict__ val, int width, int M, int2* rec)
{
    const long long n = (long long)width * M;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i / M), m = (int)(i - (long long)j * M);
        rec[i] = make_int2(col[(long long)m * width + j], __float_as_int(val[(long long)m * width + j]));
    }
}

// Same tiling as spmm_rows_kernel (TR consecutive rows x column slab, LG lanes per row).  SB > 1: a tile spans SB
// samples and a thread keeps the W records of its row in registers across them (W > 0 only).  Variants with all W record
// loads and all W gathers forced in flight, and with the tile's records staged in shared memory, were measured and
// removed: every one of them lands on the same ~60 us at cosmo layer 2 (profiles/r01_exp_spmm_limits.txt) -- the limit is
// the load/store pipe, which only the pipelined kernel (sparse_pipe.cu) relieves.
template <int PASSES, int W, int SB>
__global__ void __launch_bounds__(256)
spmm_rec_kernel(const int2* __restrict__ rec, int width, int B, int M, int chunks, int lg2, int TR, int CB,
                const float4* __restrict__ x, const float4* y, const float4* z, float4* out,
                float alpha, float beta, float gamma)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int LG = 1 << lg2, rpw = 32 >> lg2;
    const int sub = lane >> lg2, lg = lane & (LG - 1);
    const int rtiles = (M + TR - 1) / TR, ctiles = (chunks + CB - 1) / CB;
    const int bgroups = (B + SB - 1) / SB;
    const long long tiles = (long long)bgroups * rtiles * ctiles;
    const int plane = M * chunks;                                  // float4 per sample (host checks it fits in 31 bits)
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int ct = (int)(t % ctiles);
        const long long t2 = t / ctiles;
        const int rt = (int)(t2 % rtiles), bg = (int)(t2 / rtiles);
        const int coff = ct * CB + lg;
        const int mend = min(M, (rt + 1) * TR);
        bool on[PASSES];
#pragma unroll
        for (int p = 0; p < PASSES; ++p) on[p] = (coff + p * LG < chunks) && (lg + p * LG < CB);
        for (int m = rt * TR + warp * rpw + sub; m < mend; m += 8 * rpw) {
            if (W > 0) {
                int c[W > 0 ? W : 1];
                float wv[W > 0 ? W : 1];
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    const int2 r = __ldg(rec + (long long)j * M + m);
                    c[j] = r.x * chunks + coff;
                    wv[j] = __int_as_float(r.y);
                }
#pragma unroll
                for (int s = 0; s < SB; ++s) {
                    const int b = bg * SB + s;
                    if (SB > 1 && b >= B) break;
                    const float4* xb = x + (long long)b * plane;
                    float4 v[W > 0 ? W : 1][PASSES];
#pragma unroll
                    for (int j = 0; j < W; ++j) {
#pragma unroll
                        for (int p = 0; p < PASSES; ++p)
                            v[j][p] = on[p] ? __ldg(xb + c[j] + p * LG) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int p = 0; p < PASSES; ++p) {
                        if (!on[p]) continue;
                        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int j = 0; j < W; ++j) {
                            a.x = fmaf(wv[j], v[j][p].x, a.x); a.y = fmaf(wv[j], v[j][p].y, a.y);
                            a.z = fmaf(wv[j], v[j][p].z, a.z); a.w = fmaf(wv[j], v[j][p].w, a.w);
                        }
                        const long long o = (long long)b * plane + m * chunks + coff + p * LG;
                        float4 r = make_float4(a.x * alpha, a.y * alpha, a.z * alpha, a.w * alpha);
                        if (y) { const float4 t4 = y[o]; r.x = fmaf(beta, t4.x, r.x); r.y = fmaf(beta, t4.y, r.y); r.z = fmaf(beta, t4.z, r.z); r.w = fmaf(beta, t4.w, r.w); }
                        if (z) { const float4 t4 = z[o]; r.x = fmaf(gamma, t4.x, r.x); r.y = fmaf(gamma, t4.y, r.y); r.z = fmaf(gamma, t4.z, r.z); r.w = fmaf(gamma, t4.w, r.w); }
                        out[o] = r;
                    }
                }
            } else {
                const float4* xb = x + (long long)bg * plane + coff;
                float4 acc[PASSES];
#pragma unroll
                for (int p = 0; p < PASSES; ++p) acc[p] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                for (int j = 0; j < width; ++j) {
                    const int2 r = __ldg(rec + (long long)j * M + m);
                    const float wj = __int_as_float(r.y);
                    const float4* xr = xb + r.x * chunks;
#pragma unroll
                    for (int p = 0; p < PASSES; ++p) {
                        if (on[p]) {
                            const float4 v = __ldg(xr + p * LG);
                            acc[p].x = fmaf(wj, v.x, acc[p].x); acc[p].y = fmaf(wj, v.y, acc[p].y);
                            acc[p].z = fmaf(wj, v.z, acc[p].z); acc[p].w = fmaf(wj, v.w, acc[p].w);
                        }
                    }
                }
#pragma unroll
                for (int p = 0; p < PASSES; ++p) {
                    if (on[p]) {
                        const long long o = (long long)bg * plane + m * chunks + coff + p * LG;
                        float4 r = make_float4(acc[p].x * alpha, acc[p].y * alpha, acc[p].z * alpha, acc[p].w * alpha);
                        if (y) { const float4 t4 = y[o]; r.x = fmaf(beta, t4.x, r.x); r.y = fmaf(beta, t4.y, r.y); r.z = fmaf(beta, t4.z, r.z); r.w = fmaf(beta, t4.w, r.w); }
                        if (z) { const float4 t4 = z[o]; r.x = fmaf(gamma, t4.x, r.x); r.y = fmaf(gamma, t4.y, r.y); r.z = fmaf(gamma, t4.z, r.z); r.w = fmaf(gamma, t4.w, r.w); }
                        out[o] = r;
                    }
                }
            }
        }
    }
}

#define env_int(name, dflt) DSB_ENV_INT(name, dflt)

constexpr int kNotQualified = -1000;
// returns kNotQualified when the shape does not qualify (caller falls back to the row-major kernels)
static int spmm_rec_launch(const int2* rec, int width, int B, int M, int F, const float* x, const float* y, const float* z,
                           float* out, float alpha, float beta, float gamma, cudaStream_t s)
{
    if (beta == 0.f) y = nullptr;
    if (gamma == 0.f) z = nullptr;
    const uintptr_t align = (uintptr_t)x | (uintptr_t)out | (uintptr_t)y | (uintptr_t)z;
    if (F % 4 || (align & 15) || (long long)M * (F / 4) >= (1LL << 31)) return kNotQualified;
    DSB_REQUIRE(x != out, "spmm: out must not alias x");
    const int chunks = F / 4;
    int CB = chunks < 64 ? chunks : 64;
    int lg2 = 0;
    while ((1 << lg2) < CB && lg2 < 5) ++lg2;
    const int passes = (CB + (1 << lg2) - 1) >> lg2;
    int TR = 32768 / (CB * 16);
    TR = TR < 32 ? 32 : TR > 128 ? 128 : TR;
    { const int v = env_int("DSB200_SPMM_TR", 0); if (v > 0) TR = v; }
    int SB = (width == 9 || width == 7) && passes == 1 ? env_int("DSB200_SPMM_SB", 1) : 1;
    if (SB != 2 && SB != 4) SB = 1;
    while (SB > 1 && (long long)((B + SB - 1) / SB) * ((M + TR - 1) / TR) * ((chunks + CB - 1) / CB) < 8LL * kNumSMs) SB >>= 1;
    const int bgroups = (B + SB - 1) / SB;
    while (TR > 16 && (long long)bgroups * ((M + TR - 1) / TR) * ((chunks + CB - 1) / CB) < 8LL * kNumSMs) TR >>= 1;
    const long long tiles = (long long)bgroups * ((M + TR - 1) / TR) * ((chunks + CB - 1) / CB);
    const long long cap = (long long)kNumSMs * 8;
    const int grid = (int)(tiles < cap ? tiles : cap);
#define DSB_REC(PP, WW, SS) spmm_rec_kernel<PP, WW, SS><<<grid, 256, 0, s>>>(rec, width, B, M, chunks, lg2, TR, CB, \
            (const float4*)x, (const float4*)y, (const float4*)z, (float4*)out, alpha, beta, gamma)
    if (passes > 1) DSB_REC(2, 0, 1);                            // wide rows (F > 128): the unrolled form needs 128 registers
    else if (width == 9) { if (SB == 4) DSB_REC(1, 9, 4); else if (SB == 2) DSB_REC(1, 9, 2); else DSB_REC(1, 9, 1); }
    else if (width == 7) { if (SB == 4) DSB_REC(1, 7, 4); else if (SB == 2) DSB_REC(1, 7, 2); else DSB_REC(1, 7, 1); }
    else DSB_REC(1, 0, 1);
#undef DSB_REC
    DSB_LAUNCH_CHECK();
}

// one product through the packed records when they are given and the shape qualifies, else the row-major kernels
static int spmm_any(const int2* rec, const int32_t* rowptr, const int32_t* col, const float* val, int width, int B, int M, int F,
                    const float* x, const float* y, const float* z, float* out, float alpha, float beta, float gamma, cudaStream_t s)
{
    if (rec && !rowptr) {
        const int rc = spmm_rec_launch(rec, width, B, M, F, x, y, z, out, alpha, beta, gamma, s);
        if (rc != kNotQualified) return rc;
    }
    return spmm_launch(rowptr, col, val, width, B, M, F, x, y, z, out, alpha, beta, gamma, s);
}

// ------------------------------------------------------------------------------------------------------------------
// Whole recurrence of a SMALL level in one launch.  At the coarsest levels (M <= 341 vertices) a sample's 16-feature slab of
// one order is at most 21 KB: a CTA keeps X_(k-1), X_(k-2) and X_k of (sample, slab) in shared memory and runs all K - 1
// steps with a barrier in between -- K - 1 launches of a few microseconds each, all latency, become one.
//   FWD : X_1 = L X_0,  X_k = 2 L X_(k-1) - X_(k-2)                       -> basis planes
//   BWD : A_(K-1) = G_(K-1),  A_k = G_k + c_k L^T A_(k+1) - A_(k+2)      -> A_0 written over G_0   (c_k = 2, c_0 = 1)
template <int W, bool BWD>
__global__ void __launch_bounds__(256)
cheb_small_kernel(const int32_t* __restrict__ col, const float* __restrict__ val, int M, int chunks, int K,
                  const float4* x0, float4* planes, long long plane_stride)
{
    extern __shared__ float4 sp[];                              // [3][M][4]
    const int slab = blockIdx.x, b = blockIdx.y;
    const long long base = (long long)b * M * chunks + slab * 4;
    const int nitems = M * 4;
    float4* buf[3] = {sp, sp + (size_t)M * 4, sp + (size_t)2 * M * 4};
    // plane index helpers: FWD order k = 0 is x0, k >= 1 is planes[k-1]; BWD order k is planes[k] (= G_k)
    auto gptr = [&](int k) -> float4* { return BWD ? planes + (long long)k * plane_stride : (k == 0 ? const_cast<float4*>(x0) : planes + (long long)(k - 1) * plane_stride); };
    {
        const float4* src = gptr(BWD ? K - 1 : 0) + base;
        for (int i = threadIdx.x; i < nitems; i += 256) buf[0][i] = src[(long long)(i >> 2) * chunks + (i & 3)];
    }
    __syncthreads();
    int cur = 0;                                                // buf[cur] = newest order, buf[prev] = the one before
    int prev = 2;
    const int steps = K - 1;
    for (int st = 0; st < steps; ++st) {
        const int k = BWD ? K - 2 - st : st + 1;                // order produced by this step
        const int nxt = 3 - cur - prev;
        const float alpha = BWD ? (k >= 1 ? 2.f : 1.f) : (k == 1 ? 1.f : 2.f);
        const bool has_prev = st > 0;                           // X_(k-2) / A_(k+2) exists
        float4* dstg = gptr(BWD ? (k == 0 ? 0 : -1) : k);       // BWD stores only A_0
        const float4* addg = BWD ? gptr(k) + base : nullptr;    // G_k
        for (int i = threadIdx.x; i < nitems; i += 256) {
            const int r = i >> 2, c = i & 3;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < W; ++j) {
                const int cj = __ldg(col + r * W + j);
                const float w = __ldg(val + r * W + j);
                const float4 v = buf[cur][cj * 4 + c];
                acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
            }
            float4 res = make_float4(alpha * acc.x, alpha * acc.y, alpha * acc.z, alpha * acc.w);
            if (has_prev) { const float4 p2 = buf[prev][i]; res.x -= p2.x; res.y -= p2.y; res.z -= p2.z; res.w -= p2.w; }
            if (BWD) { const float4 gk = addg[(long long)r * chunks + c]; res.x += gk.x; res.y += gk.y; res.z += gk.z; res.w += gk.w; }
            buf[nxt][i] = res;
            if (!BWD || k == 0) (dstg + base)[(long long)r * chunks + c] = res;
        }
        __syncthreads();
        prev = cur;
        cur = nxt;
    }
}

// returns 1 if the level does not qualify
static int cheb_small_launch(const int32_t* col, const float* val, int width, int B, int M, int F, int K, const float* x0, float* planes,
                             bool bwd, cudaStream_t s)
{
    if (DSB_ENV_INT("DSB200_NO_SMALL", 0) || K < 2 || F % 16 || (width != 9 && width != 7)) return 1;
    const size_t smem = (size_t)3 * M * 64;
    if (smem > 64 * 1024 || M < 16) return 1;           // measured: wins at M = 256 (18.6 vs 25.9 us), loses at M = 1024 (59 vs 34 us: 64 CTAs only)
    if ((((uintptr_t)x0 | (uintptr_t)planes) & 15) != 0) return 1;
    const long long plane = (long long)B * M * (F / 4);
    dim3 grid(F / 16, B);
#define DSB_SMALL(WW, BB) do { \
        static bool cfg = false; \
        if (!cfg) { cudaError_t e = cudaFuncSetAttribute(cheb_small_kernel<WW, BB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
                    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return (int)e; } cfg = true; } \
        cheb_small_kernel<WW, BB><<<grid, 256, smem, s>>>(col, val, M, F / 4, K, (const float4*)x0, (float4*)planes, plane); } while (0)
    if (width == 9) { if (bwd) DSB_SMALL(9, true); else DSB_SMALL(9, false); }
    else { if (bwd) DSB_SMALL(7, true); else DSB_SMALL(7, false); }
#undef DSB_SMALL
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

static int cheb_basis_impl(const int2* rec, const int32_t* rowptr, const int32_t* col, const float* val, int width,
                           int B, int M, int F, int K, const float* x0, float* basis, cudaStream_t s)
{
    DSB_REQUIRE(K >= 1, "cheb_basis: K must be >= 1");
    if (K == 1) return 0;
    DSB_REQUIRE(basis, "cheb_basis: null basis");
    const long long plane = (long long)B * M * F;
    if (!rowptr) { const int rs = cheb_small_launch(col, val, width, B, M, F, K, x0, basis, false, s); if (rs != 1) return rs; }
    // X1 = L X0
    int rc = spmm_any(rec, rowptr, col, val, width, B, M, F, x0, nullptr, nullptr, basis, 1.f, 0.f, 0.f, s);
    // Xk = 2 L X(k-1) - X(k-2)
    for (int k = 2; k < K && rc == 0; ++k) {
        const float* xm1 = basis + (long long)(k - 2) * plane;
        const float* xm2 = (k == 2) ? x0 : basis + (long long)(k - 3) * plane;
        rc = spmm_any(rec, rowptr, col, val, width, B, M, F, xm1, xm2, nullptr, basis + (long long)(k - 1) * plane,
                      2.f, -1.f, 0.f, s);
    }
    return rc;
}

// Adjoint of X1 = L X0, Xk = 2 L X(k-1) - X(k-2):  with A_k = dLoss/dX_k (total),
//   A_(K-1) = G_(K-1);  A_k = G_k + 2 L^T A_(k+1) - A_(k+2)  (1 <= k <= K-2);
//   A_0 = G_0 + L^T A_1 - A_2.
static int cheb_basis_bwd_impl(const int2* recT, const int32_t* rowptrT, const int32_t* colT, const float* valT, int width,
                               int B, int M, int F, int K, float* G, cudaStream_t s)
{
    DSB_REQUIRE(K >= 1 && G, "cheb_basis_bwd: bad arguments");
    if (K == 1) return 0;
    const long long plane = (long long)B * M * F;
    if (!rowptrT) { const int rs = cheb_small_launch(colT, valT, width, B, M, F, K, nullptr, G, true, s); if (rs != 1) return rs; }
    int rc = 0;
    for (int k = K - 2; k >= 1 && rc == 0; --k) {
        float* gk = G + (long long)k * plane;
        const float* a1 = G + (long long)(k + 1) * plane;
        const float* a2 = (k + 2 <= K - 1) ? G + (long long)(k + 2) * plane : nullptr;
        rc = spmm_any(recT, rowptrT, colT, valT, width, B, M, F, a1, gk, a2, gk, 2.f, 1.f, a2 ? -1.f : 0.f, s);
    }
    if (rc) return rc;
    const float* a2 = (K > 2) ? G + 2 * plane : nullptr;
    return spmm_any(recT, rowptrT, colT, valT, width, B, M, F, G + plane, G, a2, G, 1.f, 1.f, a2 ? -1.f : 0.f, s);
}

extern "C" int dsb200_cheb_basis(const int32_t* rowptr, const int32_t* col, const float* val, int width,
                                 int B, int M, int F, int K, const float* x0, float* basis, void* stream)
{
    return cheb_basis_impl(nullptr, rowptr, col, val, width, B, M, F, K, x0, basis, (cudaStream_t)stream);
}

extern "C" int dsb200_cheb_basis_bwd(const int32_t* rowptrT, const int32_t* colT, const float* valT, int width,
                                     int B, int M, int F, int K, float* G, void* stream)
{
    return cheb_basis_bwd_impl(nullptr, rowptrT, colT, valT, width, B, M, F, K, G, (cudaStream_t)stream);
}

extern "C" int dsb200_ell_pack(const int32_t* col, const float* val, int width, int M, void* rec, void* stream)
{
    DSB_REQUIRE(col && val && rec && width > 0 && M > 0, "ell_pack: bad arguments");
    ell_pack_kernel<<<grid_for((long long)width * M, 256, 8), 256, 0, (cudaStream_t)stream>>>(col, val, width, M, (int2*)rec);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_spmm_ellp(const void* rec, const int32_t* col, const float* val, int width, int B, int M, int F,
                                const float* x, const float* y, const float* z, float* out,
                                float alpha, float beta, float gamma, void* stream)
{
    DSB_REQUIRE(B > 0 && M > 0 && F > 0, "spmm: empty matrix");
    DSB_REQUIRE(rec && col && val && x && out && width > 0, "spmm_ellp: bad arguments");
    return spmm_any((const int2*)rec, nullptr, col, val, width, B, M, F, x, y, z, out, alpha, beta, gamma, (cudaStream_t)stream);
}

extern "C" int dsb200_cheb_basis_ellp(const void* rec, const int32_t* col, const float* val, int width,
                                      int B, int M, int F, int K, const float* x0, float* basis, void* stream)
{
    DSB_REQUIRE(rec && col && val, "cheb_basis_ellp: bad arguments");
    return cheb_basis_impl((const int2*)rec, nullptr, col, val, width, B, M, F, K, x0, basis, (cudaStream_t)stream);
}

extern "C" int dsb200_cheb_basis_bwd_ellp(const void* recT, const int32_t* colT, const float* valT, int width,
                                          int B, int M, int F, int K, float* G, void* stream)
{
    DSB_REQUIRE(recT && colT && valT, "cheb_basis_bwd_ellp: bad arguments");
    return cheb_basis_bwd_impl((const int2*)recT, nullptr, colT, valT, width, B, M, F, K, G, (cudaStream_t)stream);
}
