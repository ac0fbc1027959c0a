// K1, shared-memory tiled: one or TWO steps of the Chebyshev recurrence per launch.
//   T1 = a1 * L x  + b1 * y + c1 * z            (on a row tile and its first halo ring)
//   T2 = a2 * L T1 + b2 * x + c2 * u            (on the row tile; only when out2 != NULL)
// Replaces pairs of tf.sparse_tensor_dense_matmul + `2*..-x0` (reference deepsphere/models.py:1056-1061):
// forward (X1, X2) = (L X0, 2 L X1 - X0) and (X_k, X_k+1) = (2 L X_k-1 - X_k-2, 2 L X_k - X_k-1); backward the
// adjoint pairs (A_k, A_k-1) of cheb_basis_bwd (sparse.cu).  Fusing two steps means the intermediate order is
// read from shared memory instead of HBM: 3X (+ L) bytes for two forward steps instead of 5X.
//
// A CTA owns (row tile, sample, 64-byte column slab).  Rows of a tile are consecutive, which in HEALPix nested
// (Morton) order is a compact patch of the sphere; its two halo rings are precomputed on the host from the sparsity
// pattern itself (graph.py: TilePlan), so the kernel works for any graph whose vertex order has locality.
//   phase 0  cp.async: x rows of tile + ring 1 + ring 2 -> A;  y rows of tile + ring 1 -> Bf        (coalesced 16 B)
//   phase 1  one thread per row: 9 x NCH 16-byte gathers from A (bank-conflict-free XOR swizzle)    -> Bf (in place)
//   phase 2  Bf rows of the tile -> out1 (coalesced);  second step gathers from Bf -> A (in place)  -> out2
// Per local row the host stores a 16-word record (10 fp32 values + 10 uint16 LOCAL column indices), word-major per
// tile, so that the 32 rows of a warp read every record word with one coalesced request.
#include "common.cuh"

namespace dsb200 {

constexpr int kSlots = 10;                 // neighbour slots of a row record
constexpr int kRecWords = 16;              // 10 fp32 values, 5 words of uint16 local columns, 1 pad
constexpr int kTileThreads = 256;

struct TiledArgs {
    const int4* info;          // per tile: (first record word, n1 = local rows incl. ring 1, n2 = incl. ring 2, first ext entry)
    const uint32_t* recs;      // per tile [16][n1] record words (word-major: consecutive rows are consecutive words)
    const int32_t* ext;        // global row of every local index >= nt of every tile (holes: the tile's first row)
    const float4 *x, *y, *z, *u;
    float4 *out1, *out2;
    int B, M, chunks, nslab, TR, ntiles, n1cap, n2cap;
    float a1, b1, c1, a2, b2, c2;
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Shared-memory rows are padded to an ODD number of 16-byte chunks (NCH + 1 for NCH = 2, 4).  The 16-byte bank group
// of chunk i of row r is then (STRIDE * r + i) mod 8, a bijection of r mod 8: rows whose indices differ in their low
// 3 bits -- 8 consecutive tile rows, and also their neighbours in one direction (a shifted 2 x 4 block in Morton
// order) -- land in 8 different bank groups when every lane of a quarter warp reads the same chunk.  No XOR swizzle:
// the 4 chunks of a neighbour row are read with immediate offsets from one address.
template <int NCH> struct RowStride { static constexpr int value = NCH == 1 ? 1 : NCH + 1; };

__device__ __forceinline__ void fma4(float4& acc, float w, const float4& v) {
    acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
}

// acc[i] = sum_j v[j] * S[c[j]][chunk i] for local row r: one thread owns the whole NCH-chunk row
template <int NCH, int W>
__device__ __forceinline__ void gather_row(const uint32_t* __restrict__ rec, int n1, int r, const float4* S, float4 (&acc)[NCH])
{
    float v[W];
    uint32_t cw[(W + 1) / 2];
#pragma unroll
    for (int j = 0; j < W; ++j) v[j] = __uint_as_float(__ldg(rec + j * n1 + r));
#pragma unroll
    for (int j = 0; j < (W + 1) / 2; ++j) cw[j] = __ldg(rec + (kSlots + j) * n1 + r);
#pragma unroll
    for (int i = 0; i < NCH; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < W; ++j) {
        const int cj = (j & 1) ? (int)(cw[j >> 1] >> 16) : (int)(cw[j >> 1] & 0xffffu);
        const float4* row = S + cj * RowStride<NCH>::value;
#pragma unroll
        for (int i = 0; i < NCH; ++i) fma4(acc[i], v[j], row[i]);
    }
}

template <int NCH, int W>
__global__ void __launch_bounds__(kTileThreads)
cheb_tiled_kernel(const __grid_constant__ TiledArgs a)
{
    extern __shared__ float4 smem4[];
    constexpr int RS = RowStride<NCH>::value;
    float4* A = smem4;                              // [n2cap][RS]
    float4* Bf = smem4 + (size_t)a.n2cap * RS;      // [n1cap][RS]
    const bool two = a.out2 != nullptr;

    // CTA -> (tile, sample, slab): slabs and samples of a tile are neighbours in the launch order (shared records / halos in L2)
    int t = blockIdx.x;
    const int slab = t % a.nslab; t /= a.nslab;
    const int b = t % a.B; t /= a.B;
    const int4 inf = __ldg(a.info + t);
    const int r0 = t * a.TR;
    const int nt = min(a.TR, a.M - r0);
    const int n1 = inf.y, n2 = inf.z;
    const uint32_t* rec = a.recs + inf.x;
    const int32_t* ext = a.ext + inf.w;
    const long long sbase = (long long)b * a.M * a.chunks + slab * NCH;       // chunk offset of (sample, slab)
    const int nrows1 = two ? n1 : nt;              // rows computed in step 1
    const int nload = two ? n2 : n1;               // rows of x needed

    // ---- phase 0: stage x (and y) rows, 16 bytes per thread and instruction, coalesced
    for (int idx = threadIdx.x; idx < nload * NCH; idx += kTileThreads) {
        const int r = idx / NCH, c = idx - r * NCH;
        const int gr = r < nt ? r0 + r : __ldg(ext + r - nt);
        cp_async16(A + r * RS + c, a.x + sbase + (long long)gr * a.chunks + c);
    }
    if (a.y) {
        for (int idx = threadIdx.x; idx < nrows1 * NCH; idx += kTileThreads) {
            const int r = idx / NCH, c = idx - r * NCH;
            const int gr = r < nt ? r0 + r : __ldg(ext + r - nt);
            cp_async16(Bf + r * RS + c, a.y + sbase + (long long)gr * a.chunks + c);
        }
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- phase 1: T1 on the tile (+ ring 1), one thread per row
    for (int r = threadIdx.x; r < nrows1; r += kTileThreads) {
        float4 acc[NCH];
        gather_row<NCH, W>(rec, n1, r, A, acc);
        float4* mine = Bf + r * RS;
        const float4* zr = nullptr;
        if (a.z) zr = a.z + sbase + (long long)(r < nt ? r0 + r : __ldg(ext + r - nt)) * a.chunks;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            float4 res = make_float4(a.a1 * acc[i].x, a.a1 * acc[i].y, a.a1 * acc[i].z, a.a1 * acc[i].w);
            if (a.y) fma4(res, a.b1, mine[i]);
            if (a.z) fma4(res, a.c1, __ldg(zr + i));
            mine[i] = res;
        }
    }
    __syncthreads();

    // ---- phase 2: T1 of the tile rows -> out1; second step from Bf into A (in place), then -> out2
    if (a.out1) {
        for (int idx = threadIdx.x; idx < nt * NCH; idx += kTileThreads) {
            const int r = idx / NCH, c = idx - r * NCH;
            a.out1[sbase + (long long)(r0 + r) * a.chunks + c] = Bf[r * RS + c];
        }
    }
    if (!two) return;
    for (int r = threadIdx.x; r < nt; r += kTileThreads) {
        float4 acc[NCH];
        gather_row<NCH, W>(rec, n1, r, Bf, acc);
        float4* mine = A + r * RS;
        const float4* ur = a.u ? a.u + sbase + (long long)(r0 + r) * a.chunks : nullptr;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            float4 res = make_float4(a.a2 * acc[i].x, a.a2 * acc[i].y, a.a2 * acc[i].z, a.a2 * acc[i].w);
            fma4(res, a.b2, mine[i]);
            if (a.u) fma4(res, a.c2, __ldg(ur + i));
            mine[i] = res;
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < nt * NCH; idx += kTileThreads) {
        const int r = idx / NCH, c = idx - r * NCH;
        a.out2[sbase + (long long)(r0 + r) * a.chunks + c] = A[r * RS + c];
    }
}

template <int NCH, int W>
static int launch_tiled_w(const TiledArgs& a, size_t smem, cudaStream_t s)
{
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(cheb_tiled_kernel<NCH, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    const long long grid = (long long)a.ntiles * a.B * a.nslab;
    DSB_REQUIRE(grid > 0 && grid <= 0x7fffffffLL, "cheb_tiled: grid too large");
    cheb_tiled_kernel<NCH, W><<<(unsigned)grid, kTileThreads, smem, s>>>(a);
    DSB_LAUNCH_CHECK();
}

template <int NCH>
static int launch_tiled(const TiledArgs& a, int width, size_t smem, cudaStream_t s)
{
    if (width <= 7) return launch_tiled_w<NCH, 7>(a, smem, s);
    if (width <= 9) return launch_tiled_w<NCH, 9>(a, smem, s);
    return launch_tiled_w<NCH, kSlots>(a, smem, s);
}

static int tiled_step(const int32_t* info, const void* recs, const int32_t* ext, int ntiles, int TR, int n1cap, int n2cap,
                      int width, int B, int M, int F, const float* x, const float* y, const float* z, const float* u,
                      float* out1, float* out2, float a1, float b1, float c1, float a2, float b2, float c2, cudaStream_t s)
{
    DSB_REQUIRE(info && recs && ext && x && ntiles > 0 && TR > 0, "cheb_tiled: null plan or input");
    DSB_REQUIRE(B > 0 && M > 0 && F > 0 && F % 4 == 0, "cheb_tiled: F must be a positive multiple of 4");
    DSB_REQUIRE(width >= 1 && width <= kSlots, "cheb_tiled: at most 10 entries per row");
    DSB_REQUIRE(out1 || out2, "cheb_tiled: no output");
    DSB_REQUIRE(x != out1 && x != out2, "cheb_tiled: outputs must not alias x (halo rows of x are read by other tiles)");
    DSB_REQUIRE(!out2 || !out1 || (y != out1 && z != out1), "cheb_tiled: out1 must not alias y / z in a two-step launch (ring-1 reads)");
    const uintptr_t al = (uintptr_t)x | (uintptr_t)y | (uintptr_t)z | (uintptr_t)u | (uintptr_t)out1 | (uintptr_t)out2;
    DSB_REQUIRE((al & 15) == 0, "cheb_tiled: tensors must be 16-byte aligned");
    if (b1 == 0.f) y = nullptr;
    if (c1 == 0.f) z = nullptr;
    if (c2 == 0.f) u = nullptr;
    const int chunks = F / 4;
    const int nch = chunks % 4 == 0 ? 4 : chunks % 2 == 0 ? 2 : 1;
    TiledArgs a;
    a.info = (const int4*)info; a.recs = (const uint32_t*)recs; a.ext = ext;
    a.x = (const float4*)x; a.y = (const float4*)y; a.z = (const float4*)z; a.u = (const float4*)u;
    a.out1 = (float4*)out1; a.out2 = (float4*)out2;
    a.B = B; a.M = M; a.chunks = chunks; a.nslab = chunks / nch; a.TR = TR; a.ntiles = ntiles;
    a.n1cap = out2 ? n1cap : (TR + 7) / 8 * 8; a.n2cap = out2 ? n2cap : n1cap;
    a.a1 = a1; a.b1 = b1; a.c1 = c1; a.a2 = a2; a.b2 = b2; a.c2 = c2;
    const size_t smem = (size_t)(a.n1cap + a.n2cap) * (nch == 1 ? 1 : nch + 1) * 16;
    DSB_REQUIRE(smem <= 200 * 1024, "cheb_tiled: tile + halo does not fit shared memory");
    if (nch == 4) return launch_tiled<4>(a, width, smem, s);
    if (nch == 2) return launch_tiled<2>(a, width, smem, s);
    return launch_tiled<1>(a, width, smem, s);
}

}  // namespace dsb200

using namespace dsb200;

extern "C" int dsb200_cheb_tiled_step(const int32_t* tile_info, const void* tile_recs, const int32_t* tile_ext,
                                      int ntiles, int TR, int n1cap, int n2cap, int width, int B, int M, int F,
                                      const float* x, const float* y, const float* z, const float* u,
                                      float* out1, float* out2, float a1, float b1, float c1, float a2, float b2, float c2,
                                      void* stream)
{
    return tiled_step(tile_info, tile_recs, tile_ext, ntiles, TR, n1cap, n2cap, width, B, M, F, x, y, z, u, out1, out2,
                      a1, b1, c1, a2, b2, c2, (cudaStream_t)stream);
}

// Orders 1..K-1 of the Chebyshev basis, two orders per launch (basis = [K-1][B, M, F]).
extern "C" int dsb200_cheb_basis_tiled(const int32_t* tile_info, const void* tile_recs, const int32_t* tile_ext,
                                       int ntiles, int TR, int n1cap, int n2cap, int width, int B, int M, int F, int K,
                                       const float* x0, float* basis, void* stream)
{
    DSB_REQUIRE(K >= 1, "cheb_basis_tiled: K must be >= 1");
    if (K == 1) return 0;
    DSB_REQUIRE(basis && x0, "cheb_basis_tiled: null tensor");
    const long long plane = (long long)B * M * F;
    cudaStream_t s = (cudaStream_t)stream;
    auto X = [&](int k) -> float* { return k == 0 ? const_cast<float*>(x0) : basis + (long long)(k - 1) * plane; };
    int rc = 0;
    for (int k = 1; k <= K - 1 && rc == 0;) {
        const float a1 = k == 1 ? 1.f : 2.f, b1 = k == 1 ? 0.f : -1.f;
        const float* y = k == 1 ? nullptr : X(k - 2);
        if (k + 1 <= K - 1) {
            rc = tiled_step(tile_info, tile_recs, tile_ext, ntiles, TR, n1cap, n2cap, width, B, M, F, X(k - 1), y, nullptr, nullptr,
                            X(k), X(k + 1), a1, b1, 0.f, 2.f, -1.f, 0.f, s);
            k += 2;
        } else {
            rc = tiled_step(tile_info, tile_recs, tile_ext, ntiles, TR, n1cap, n2cap, width, B, M, F, X(k - 1), y, nullptr, nullptr,
                            X(k), nullptr, a1, b1, 0.f, 0.f, 0.f, 0.f, s);
            k += 1;
        }
    }
    return rc;
}

// Adjoint recurrence (see dsb200_cheb_basis_bwd), two orders per launch, with the plan of L^T:
//   A_(K-1) = G_(K-1);  A_k = G_k + c_k L^T A_(k+1) - A_(k+2),  c_k = 2 (k >= 1), 1 (k = 0).
// G [K][B, M, F] is read only; S [K][B, M, F] is scratch for the A_k; the result A_0 is S[0].
extern "C" int dsb200_cheb_basis_bwd_tiled(const int32_t* tile_info, const void* tile_recs, const int32_t* tile_ext,
                                           int ntiles, int TR, int n1cap, int n2cap, int width, int B, int M, int F, int K,
                                           const float* G, float* S, void* stream)
{
    DSB_REQUIRE(K >= 2 && G && S, "cheb_basis_bwd_tiled: bad arguments");
    const long long plane = (long long)B * M * F;
    cudaStream_t s = (cudaStream_t)stream;
    auto Gk = [&](int k) -> const float* { return G + (long long)k * plane; };
    auto Ak = [&](int k) -> float* { return k == K - 1 ? const_cast<float*>(Gk(K - 1)) : S + (long long)k * plane; };
    int rc = 0;
    for (int k = K - 2; k >= 0 && rc == 0;) {
        const float ck = k >= 1 ? 2.f : 1.f;
        const float* z = (k + 2 <= K - 1) ? Ak(k + 2) : nullptr;
        if (k >= 1) {
            const float ck1 = (k - 1) >= 1 ? 2.f : 1.f;
            // A_k is needed later only if another step follows the pair
            float* o1 = (k - 1 > 0) ? Ak(k) : nullptr;
            rc = tiled_step(tile_info, tile_recs, tile_ext, ntiles, TR, n1cap, n2cap, width, B, M, F, Ak(k + 1), Gk(k), z, Gk(k - 1),
                            o1, Ak(k - 1), ck, 1.f, z ? -1.f : 0.f, ck1, -1.f, 1.f, s);
            k -= 2;
        } else {
            rc = tiled_step(tile_info, tile_recs, tile_ext, ntiles, TR, n1cap, n2cap, width, B, M, F, Ak(k + 1), Gk(k), z, nullptr,
                            Ak(k), nullptr, ck, 1.f, z ? -1.f : 0.f, 0.f, 0.f, 0.f, s);
            k -= 1;
        }
    }
    return rc;
}
