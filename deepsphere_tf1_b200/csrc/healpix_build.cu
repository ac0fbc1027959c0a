// Device-side construction of the HEALPix 8-neighbour graph and its rescaled Laplacian in ELL form (replaces
// the host path reference deepsphere/utils.py:25-133 (healpix_weightmatrix), 231-242 (build_laplacian), 379-385 (rescale_L); the
// two healpy calls pix2vec / get_all_neighbours of utils.py:56,63 are restated from the HEALPix NESTED scheme).
// The arithmetic follows the numpy path operation by operation with explicitly rounded intrinsics (no FMA contraction), so that the
// float32 results have the same bits; the three array operations whose rounding belongs to numpy's libm (mean, exp, power) are
// left to the host between the kernels (graph_build.py):
//   hpx_coords      nested pixel -> unit vector, double arithmetic, cast to float32                       (utils.py:56-58)
//   hpx_neighbours  8 neighbours (SW, W, NW, N, NE, E, SE, S; -1 = none or outside the vertex set) and the squared chord
//                   distances ((dx^2 + dy^2) + dz^2 in float32                                              (utils.py:63-103)
//   [host: kernel width = mean of the distances, weights = exp(-d^2 / (2 width))]                          (utils.py:107-111)
//   hpx_sort_degree entries of every row in ascending column order (scipy's canonical CSR) and the degree summed in the order of
//                   scipy's W.sum(1) (np.add.reduceat: first entry + sequential sum of the rest)
//   [host: d^-1/2 = power(d, -0.5)]                                                                        (utils.py:238)
//   hpx_laplacian   I - (d_i^-1/2 w_ij) d_j^-1/2 as ELL rows (diagonal at its sorted place, padding = (row, 0))
//   ell_rescale     L * recip - I, recip = float32(1 / (lmax / 2)): scipy evaluates `L /= s` as `data *= 1.0 / s` (utils.py:383-385)
//   ell_transpose   values of L^T on the same pattern ((d_i w) d_j is symmetric only to one ulp)
#include "common.cuh"

namespace dsb200 {
namespace {

__constant__ int c_jrll[12] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4};
__constant__ int c_jpll[12] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};
__constant__ int c_dx[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
__constant__ int c_dy[8] = {0, 1, 1, 1, 0, -1, -1, -1};
__constant__ int c_face_nb[9][12] = {
    {8, 9, 10, 11, -1, -1, -1, -1, 10, 11, 8, 9}, {5, 6, 7, 4, 8, 9, 10, 11, 9, 10, 11, 8},
    {-1, -1, -1, -1, 5, 6, 7, 4, -1, -1, -1, -1}, {4, 5, 6, 7, 11, 8, 9, 10, 11, 8, 9, 10},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11},       {1, 2, 3, 0, 0, 1, 2, 3, 5, 6, 7, 4},
    {-1, -1, -1, -1, 7, 4, 5, 6, -1, -1, -1, -1}, {3, 0, 1, 2, 3, 0, 1, 2, 4, 5, 6, 7},
    {2, 3, 0, 1, -1, -1, -1, -1, 0, 1, 2, 3}};
__constant__ int c_swap[9][3] = {{0, 0, 3}, {0, 0, 6}, {0, 0, 0}, {0, 0, 5}, {0, 0, 0}, {5, 0, 0}, {0, 0, 0}, {6, 0, 0}, {3, 0, 0}};

__device__ __forceinline__ unsigned compact_even(unsigned long long v)
{
    v &= 0x5555555555555555ull;
    v = (v | (v >> 1)) & 0x3333333333333333ull;
    v = (v | (v >> 2)) & 0x0F0F0F0F0F0F0F0Full;
    v = (v | (v >> 4)) & 0x00FF00FF00FF00FFull;
    v = (v | (v >> 8)) & 0x0000FFFF0000FFFFull;
    v = (v | (v >> 16)) & 0x00000000FFFFFFFFull;
    return (unsigned)v;
}
__device__ __forceinline__ unsigned long long spread(unsigned long long v)
{
    v &= 0x00000000FFFFFFFFull;
    v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
    v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
    v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
}

__global__ void __launch_bounds__(256)
hpx_coords_kernel(int nside, int order, const long long* __restrict__ ids, long long M, float* __restrict__ coords)
{
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= M) return;
    const long long pix = ids ? ids[v] : v;
    const int face = (int)(pix >> (2 * order));
    const unsigned long long r = (unsigned long long)pix & ((1ull << (2 * order)) - 1);
    const long long ix = compact_even(r), iy = compact_even(r >> 1);
    const double npix = 12.0 * (double)nside * (double)nside;
    const double fact2 = 4.0 / npix;
    const double fact1 = __dmul_rn((double)(nside << 1), fact2);
    const long long jr = (long long)c_jrll[face] * nside - ix - iy - 1;
    const bool north = jr < nside, south = jr > 3LL * nside;
    const long long nr = north ? jr : (south ? 4LL * nside - jr : nside);
    const double cap = __dmul_rn((double)(nr * nr), fact2);
    const double z = north ? __dsub_rn(1.0, cap) : (south ? __dsub_rn(cap, 1.0) : __dmul_rn((double)(2LL * nside - jr), fact1));
    long long t = (long long)c_jpll[face] * nr + ix - iy;
    if (t < 0) t += 8 * nr;
    const double halfpi = 0.5 * 3.141592653589793;
    const double phi = (nr == nside) ? __dmul_rn(__dmul_rn(0.75 * halfpi, (double)t), fact1)
                                     : __ddiv_rn(__dmul_rn(0.5 * halfpi, (double)t), (double)nr);
    const bool near_pole = (north && z > 0.99) || (south && z < -0.99);
    const double sth = near_pole ? sqrt(__dmul_rn(cap, __dsub_rn(2.0, cap))) : sqrt(__dmul_rn(__dsub_rn(1.0, z), __dadd_rn(1.0, z)));
    coords[3 * v + 0] = (float)__dmul_rn(sth, cos(phi));
    coords[3 * v + 1] = (float)__dmul_rn(sth, sin(phi));
    coords[3 * v + 2] = (float)z;
}

// vertex of a nested pixel: `inv` = inverse map over the whole sphere (arbitrary vertex sets) or NULL (vertices = pixels 0..M-1)
__device__ __forceinline__ int vertex_of(long long pix, const int* __restrict__ inv, long long M)
{
    if (pix < 0) return -1;
    if (inv) return inv[pix];
    return pix < M ? (int)pix : -1;
}

__global__ void __launch_bounds__(256)
hpx_neighbours_kernel(int nside, int order, const long long* __restrict__ ids, const int* __restrict__ inv, long long M,
                      const float* __restrict__ coords, int* __restrict__ nb, float* __restrict__ d2)
{
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= M) return;
    const long long pix = ids ? ids[v] : v;
    const int face = (int)(pix >> (2 * order));
    const unsigned long long r = (unsigned long long)pix & ((1ull << (2 * order)) - 1);
    const int ix = (int)compact_even(r), iy = (int)compact_even(r >> 1);
    const float cx = coords[3 * v], cy = coords[3 * v + 1], cz = coords[3 * v + 2];
#pragma unroll
    for (int d = 0; d < 8; ++d) {
        int x = ix + c_dx[d], y = iy + c_dy[d], code = 4;
        if (x < 0) { x += nside; code -= 1; } else if (x >= nside) { x -= nside; code += 1; }
        if (y < 0) { y += nside; code -= 3; } else if (y >= nside) { y -= nside; code += 3; }
        const int f = c_face_nb[code][face];
        long long np_ = -1;
        if (f >= 0) {
            const int bits = c_swap[code][face >> 2];
            if (bits & 1) x = nside - x - 1;
            if (bits & 2) y = nside - y - 1;
            if (bits & 4) { const int t = x; x = y; y = t; }
            np_ = ((long long)f << (2 * order)) + (long long)spread((unsigned)x) + ((long long)spread((unsigned)y) << 1);
        }
        const int u = vertex_of(np_, inv, M);
        nb[8 * v + d] = u;
        float dist = 0.f;
        if (u >= 0) {
            const float ex = __fsub_rn(cx, coords[3 * (long long)u]), ey = __fsub_rn(cy, coords[3 * (long long)u + 1]),
                        ez = __fsub_rn(cz, coords[3 * (long long)u + 2]);
            dist = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
        }
        d2[8 * v + d] = dist;
    }
}

__global__ void __launch_bounds__(256)
hpx_sort_degree_kernel(const int* __restrict__ nb, const float* __restrict__ w, long long M, int* __restrict__ scol,
                       float* __restrict__ sw, int* __restrict__ cnt, float* __restrict__ deg)
{
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= M) return;
    int c[8];
    float x[8];
    int n = 0;
    for (int d = 0; d < 8; ++d) {
        const int u = nb[8 * v + d];
        if (u < 0) continue;
        const float wv = w[8 * v + d];
        int j = n++;
        while (j > 0 && c[j - 1] > u) { c[j] = c[j - 1]; x[j] = x[j - 1]; --j; }      // insertion sort, ascending column
        c[j] = u; x[j] = wv;
    }
    // duplicates (a pixel reached in two directions: nside 1) are summed like scipy's sum_duplicates
    int m = 0;
    for (int j = 0; j < n; ++j) {
        if (m > 0 && c[m - 1] == c[j]) x[m - 1] = __fadd_rn(x[m - 1], x[j]);
        else { c[m] = c[j]; x[m] = x[j]; ++m; }
    }
    // W.sum(1) of scipy's CSR = np.add.reduceat over the row: the first entry plus the sequential sum of the others
    float s = 0.f;
    if (m > 0) {
        s = x[0];
        if (m > 1) {
            float tail = x[1];
            for (int j = 2; j < m; ++j) tail = __fadd_rn(tail, x[j]);
            s = __fadd_rn(s, tail);
        }
    }
    for (int j = 0; j < 8; ++j) { scol[8 * v + j] = j < m ? c[j] : -1; sw[8 * v + j] = j < m ? x[j] : 0.f; }
    cnt[v] = m;
    deg[v] = s;
}

__global__ void __launch_bounds__(256)
hpx_laplacian_kernel(const int* __restrict__ scol, const float* __restrict__ sw, const int* __restrict__ cnt,
                     const float* __restrict__ d12, long long M, int width, int normalized, const float* __restrict__ deg,
                     int* __restrict__ col, float* __restrict__ val)
{
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= M) return;
    const int m = cnt[v];
    const float di = normalized ? d12[v] : 0.f;
    int o = 0;
    bool diag_done = false;
    for (int j = 0; j <= m; ++j) {
        const int u = j < m ? scol[8 * v + j] : 0x7fffffff;
        if (!diag_done && u > v) {                            // the diagonal entry at its sorted place
            col[v * width + o] = (int)v;
            val[v * width + o] = normalized ? 1.f : deg[v];
            ++o;
            diag_done = true;
        }
        if (j < m) {
            const float wv = sw[8 * v + j];
            col[v * width + o] = u;
            val[v * width + o] = normalized ? -__fmul_rn(__fmul_rn(di, wv), d12[u]) : -wv;
            ++o;
        }
    }
    for (; o < width; ++o) { col[v * width + o] = (int)v; val[v * width + o] = 0.f; }
}

__global__ void __launch_bounds__(256)
ell_rescale_kernel(const int* __restrict__ col, float* __restrict__ val, long long M, int width, float recip)
{
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= M) return;
    bool diag_done = false;
    for (int j = 0; j < width; ++j) {
        float x = __fmul_rn(val[v * width + j], recip);
        if (!diag_done && col[v * width + j] == v) { x = __fsub_rn(x, 1.f); diag_done = true; }   // first (row, .) slot = the diagonal
        else if (col[v * width + j] == v) x = 0.f;                                                 // padding
        val[v * width + j] = x;
    }
}

__global__ void __launch_bounds__(256)
ell_transpose_kernel(const int* __restrict__ col, const float* __restrict__ val, long long M, int width, float* __restrict__ valT,
                     int* __restrict__ missing)
{
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= M) return;
    bool diag_done = false;
    for (int j = 0; j < width; ++j) {
        const int u = col[v * width + j];
        float x = val[v * width + j];
        if (u == v) {
            if (diag_done) x = 0.f;
            diag_done = true;
        } else {
            bool found = false;
            for (int k = 0; k < width; ++k)
                if (col[(long long)u * width + k] == (int)v) { x = val[(long long)u * width + k]; found = true; break; }
            if (!found) { atomicAdd(missing, 1); x = 0.f; }
        }
        valT[v * width + j] = x;
    }
}

}  // namespace
}  // namespace dsb200

using namespace dsb200;
#define HPX_GRID(M) (unsigned)(((M) + 255) / 256), 256, 0, (cudaStream_t)stream

extern "C" int dsb200_hpx_coords(int nside, const long long* ids, long long M, float* coords, void* stream)
{
    DSB_REQUIRE(nside > 0 && (nside & (nside - 1)) == 0 && M > 0 && coords, "hpx_coords: nside must be a power of 2");
    int order = 0;
    while ((1 << order) < nside) ++order;
    hpx_coords_kernel<<<HPX_GRID(M)>>>(nside, order, ids, M, coords);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_hpx_neighbours(int nside, const long long* ids, const int32_t* inverse, long long M, const float* coords,
                                     int32_t* nb, float* d2, void* stream)
{
    DSB_REQUIRE(nside > 0 && (nside & (nside - 1)) == 0 && M > 0 && coords && nb && d2, "hpx_neighbours: bad arguments");
    int order = 0;
    while ((1 << order) < nside) ++order;
    hpx_neighbours_kernel<<<HPX_GRID(M)>>>(nside, order, ids, inverse, M, coords, nb, d2);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_hpx_sort_degree(const int32_t* nb, const float* w, long long M, int32_t* scol, float* sw, int32_t* cnt,
                                      float* deg, void* stream)
{
    DSB_REQUIRE(nb && w && scol && sw && cnt && deg && M > 0, "hpx_sort_degree: bad arguments");
    hpx_sort_degree_kernel<<<HPX_GRID(M)>>>(nb, w, M, scol, sw, cnt, deg);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_hpx_laplacian(const int32_t* scol, const float* sw, const int32_t* cnt, const float* d12, const float* deg,
                                    long long M, int width, int normalized, int32_t* col, float* val, void* stream)
{
    DSB_REQUIRE(scol && sw && cnt && col && val && M > 0 && width >= 2 && width <= 9, "hpx_laplacian: bad arguments");
    DSB_REQUIRE(normalized ? d12 != nullptr : deg != nullptr, "hpx_laplacian: degrees missing");
    hpx_laplacian_kernel<<<HPX_GRID(M)>>>(scol, sw, cnt, d12, M, width, normalized, deg, col, val);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_ell_rescale(const int32_t* col, float* val, long long M, int width, float recip, void* stream)
{
    DSB_REQUIRE(col && val && M > 0 && width > 0 && recip > 0.f, "ell_rescale: bad arguments");
    ell_rescale_kernel<<<HPX_GRID(M)>>>(col, val, M, width, recip);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_ell_transpose(const int32_t* col, const float* val, long long M, int width, float* valT, int32_t* missing,
                                    void* stream)
{
    DSB_REQUIRE(col && val && valT && missing && M > 0 && width > 0, "ell_transpose: bad arguments");
    ell_transpose_kernel<<<HPX_GRID(M)>>>(col, val, M, width, valT, missing);
    DSB_LAUNCH_CHECK();
}
