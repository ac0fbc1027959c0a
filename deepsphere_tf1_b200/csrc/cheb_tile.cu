// K1 fused: ALL K - 1 sparse products of a ChebConv layer in one launch, per tile, out of shared memory.
//   forward :  X_1 = L X_0,  X_k = 2 L X_(k-1) - X_(k-2)            (cgcnn.chebyshev5, reference deepsphere/models.py:1049-1061;
//              with `monomial`: X_k = L X_(k-1), models.py:1094-1103)  -> planes 1..K-1 written once, X_0 read once (+ halo)
//   backward:  A_(K-1) = G_(K-1),  A_k = G_k + c_k L^T A_(k+1) - A_(k+2),  dX = A_0   (TF autodiff of the same lines, Clenshaw
//              form)                                                    -> the K planes of G read once (+ halo), dX written once
// Per-step launches move 3 X bytes per step (11 X forward, 16 X backward per layer); here the intermediate orders never leave
// the SM.  What makes that pay: HEALPix NESTED order is Morton order inside a base face, so 256 consecutive vertices are a
// 16 x 16 square of pixels and the 8-neighbour Laplacian is a 3 x 3 stencil with per-pixel weights.  A tile is staged with an
// S-ring halo (S = K - 1) as a flat (16 + 2S)^2 image in shared memory (graph.py: StencilPlan gives, per tile, the vertex at
// every image position and the 3 x 3 weights at every computed position, verified against the matrix); step s is computed on
// the image shrunk by s + 1 rings.  A warp owns a strip of 8 pixel columns x 16 features and slides down the rows keeping the
// 3 x 3 window in registers: 3 new 16-byte shared-memory reads per output instead of 9 gathers (the per-step kernels are bound
// by the load/store pipe of the SM, not by HBM: profiles/r01_exp_spmm_limits.txt), the weights arrive as broadcast reads.
// Two image buffers: X_k overwrites X_(k-2) in place (read by the same thread that writes it).
// Work item = (tile, sample, chunk of 16 features); a CTA takes a contiguous range of items (tile-major, so the tile's plan
// is loaded once for all its samples); 2 CTAs per SM overlap one's staging with the other's arithmetic.
#include "common.cuh"

namespace dsb200 {
namespace {

constexpr int kT = 16;                 // tile side (pixels)
constexpr int kWarps = 6;
constexpr int kThreads = kWarps * 32;
constexpr int kWF = 12;                // floats per weight record (9 used)
constexpr int kMaxS = 7;

struct TileArgs {
    const int* p2v;        // [nt][P * P]
    const float* wts;      // [nt][Q * Q][12],  Q = P - 2
    const float* x;        // forward: X_0 [B][M][F]; backward: G [S + 1][B][M][F]
    float* out;            // forward: basis [S][B][M][F]; backward: dX [B][M][F]
    int nt, B, M, F, S, P, NC, mono;
    long long plane;       // B * M * F
    long long nitems;      // nt * B * NC
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ float4 fma4(float w, float4 v, float4 a) {
    a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
    return a;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }

// One output row of a strip: the 3 x 3 window (rows ra = y - 1, rb = y, rc = y + 1 of the input image, columns x - 1 .. x + 1)
// times the 9 weights of the pixel.  `c*` are loaded here and become the `b*` of the next row (the caller rotates by name).
#define DSB_TILE_ROW(a0, a1, a2, b0, b1, b2, c0, c1, c2)                                                         \
    {                                                                                                            \
        c0 = pin[4 * (P - 1)]; c1 = pin[4 * P]; c2 = pin[4 * (P + 1)];                                           \
        const float4 w0 = pw[0], w1 = pw[1];                                                                     \
        const float w8 = reinterpret_cast<const float*>(pw + 2)[0];                                              \
        const int vcur = v;                                                                                      \
        float4 ecur;                                                                                             \
        if (BWD) ecur = e;                                                                                       \
        if (y + 1 < y1) {                                                                                        \
            v = pv[P];                                                                                           \
            if (BWD) { e = make_float4(0.f, 0.f, 0.f, 0.f); if (v >= 0) e = ldg4(ext + (size_t)v * F); }         \
        }                                                                                                        \
        float4 acc;                                                                                              \
        acc.x = w0.x * a0.x; acc.y = w0.x * a0.y; acc.z = w0.x * a0.z; acc.w = w0.x * a0.w;                      \
        acc = fma4(w0.y, a1, acc); acc = fma4(w0.z, a2, acc);                                                    \
        acc = fma4(w0.w, b0, acc); acc = fma4(w1.x, b1, acc); acc = fma4(w1.y, b2, acc);                         \
        acc = fma4(w1.z, c0, acc); acc = fma4(w1.w, c1, acc); acc = fma4(w8, c2, acc);                           \
        float4 res;                                                                                              \
        if (has_prev) {                                                                                          \
            const float4 pvv = pob[0];                                                                           \
            res.x = fmaf(alpha, acc.x, -pvv.x); res.y = fmaf(alpha, acc.y, -pvv.y);                              \
            res.z = fmaf(alpha, acc.z, -pvv.z); res.w = fmaf(alpha, acc.w, -pvv.w);                              \
        } else {                                                                                                 \
            res.x = alpha * acc.x; res.y = alpha * acc.y; res.z = alpha * acc.z; res.w = alpha * acc.w;          \
        }                                                                                                        \
        if (BWD) { res.x += ecur.x; res.y += ecur.y; res.z += ecur.z; res.w += ecur.w; }                         \
        if (active && !last) pob[0] = res;                                                                       \
        if (active && gout != nullptr && xin && y >= S && y < S + kT && vcur >= 0) st4(gout + (size_t)vcur * F, res); \
        pin += 4 * P; pw += 3 * Q; pob += 4 * P; pv += P; ++y;                                                   \
    }

// SS > 0: K - 1 known at compile time (image side P = 16 + 2 SS); SS == 0: run-time.
template <bool BWD, int SS>
__global__ void __launch_bounds__(kThreads, 2) cheb_tile_kernel(const TileArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = SS > 0 ? SS : a.S;
    const int P = kT + 2 * S, Q = P - 2, F = a.F;
    float4* buf0 = reinterpret_cast<float4*>(smem_raw);
    float4* buf1 = buf0 + P * P * 4;
    float4* w4 = buf1 + P * P * 4;
    int* p2v = reinterpret_cast<int*>(w4 + Q * Q * 3);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lc = lane >> 2, q = lane & 3;

    const long long i0 = a.nitems * blockIdx.x / gridDim.x, i1 = a.nitems * (blockIdx.x + 1) / gridDim.x;
    const int per_tile = a.B * a.NC;
    int cur_tile = -1;
    for (long long item = i0; item < i1; ++item) {
        const int t = (int)(item / per_tile), r = (int)(item % per_tile), b = r / a.NC, c = r % a.NC;
        __syncthreads();                                       // the previous item's last step has read both images
        if (t != cur_tile) {
            const int* gp = a.p2v + (size_t)t * P * P;
            for (int i = tid; i < P * P / 4; i += kThreads) cp_async16(p2v + 4 * i, gp + 4 * i);
            const float* gw = a.wts + (size_t)t * Q * Q * kWF;
            for (int i = tid; i < Q * Q * 3; i += kThreads) cp_async16(w4 + i, gw + 4 * i);
            cp_async_wait_all();
            __syncthreads();
            cur_tile = t;
        }
        const size_t sample = ((size_t)b * a.M) * F + (size_t)c * 16 + q * 4;      // + vertex * F: this lane's 4 features
        {   // stage the first image: X_0 (forward) / G_(K-1) (backward) at every position that shows a vertex, zeros elsewhere
            const float* src = a.x + (BWD ? (size_t)S * a.plane : 0) + ((size_t)b * a.M) * F + (size_t)c * 16;
            for (int i = tid; i < P * P * 4; i += kThreads) {
                const int v = p2v[i >> 2];
                if (v >= 0) cp_async16(buf0 + i, src + (size_t)v * F + (i & 3) * 4);
                else buf0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (BWD) {
                // the G_k rows the steps will add: into L2 now, so that their loads inside the steps are L2 hits
                // (64-byte pieces: one request per pixel covers this item's chunk)
                for (int i = tid; i < Q * Q; i += kThreads) {
                    const int py = i / Q + 1, px = i % Q + 1;
                    const int v = p2v[py * P + px];
                    if (v >= 0) {
                        const float* g0 = a.x + ((size_t)b * a.M + v) * F + (size_t)c * 16;
                        for (int k = 0; k < S; ++k) {
                            const int ring = S - 1 - k;                            // G_k is read on the image shrunk by S - k rings
                            if (py > ring && py < P - 1 - ring && px > ring && px < P - 1 - ring)
                                prefetch_l2(g0 + (size_t)k * a.plane);
                        }
                    }
                }
            }
            cp_async_wait_all();
            __syncthreads();
        }
#pragma unroll 1
        for (int s = 0; s < S; ++s) {
            const int lo = s + 1, n = P - 2 * lo;
            const int nstrips = (n + 7) >> 3, nseg = kWarps / nstrips;
            const float4* in = (s & 1) ? buf1 : buf0;
            float4* ob = (s & 1) ? buf0 : buf1;
            const bool last = s == S - 1;
            float alpha;
            bool has_prev;
            const float* ext = nullptr;
            float* gout = nullptr;
            if (BWD) {
                const int k = S - 1 - s;
                alpha = (k >= 1 && !a.mono) ? 2.f : 1.f;
                has_prev = s >= 1 && !a.mono;
                ext = a.x + (size_t)k * a.plane + sample;
                if (last) gout = a.out + sample;
            } else {
                alpha = (s == 0 || a.mono) ? 1.f : 2.f;
                has_prev = !(s == 0 || a.mono);
                gout = a.out + (size_t)s * a.plane + sample;
            }
            if (warp < nstrips * nseg) {
                const int strip = warp % nstrips, seg = warp / nstrips;
                const int y0 = lo + seg * n / nseg, y1 = lo + (seg + 1) * n / nseg;
                const int x = lo + strip * 8 + lc;
                const bool active = x < lo + n;
                const int xc = active ? x : lo + n - 1;                        // idle lanes repeat the last column (no stores)
                const bool xin = x >= S && x < S + kT;
                const float4* pin = in + ((y0 - 1) * P + xc) * 4 + q;          // row y - 1 of the window
                float4 a0 = pin[-4], a1 = pin[0], a2 = pin[4];
                pin += 4 * P;                                                  // row y
                float4 b0 = pin[-4], b1 = pin[0], b2 = pin[4];
                float4 c0, c1, c2;
                const float4* pw = w4 + ((y0 - 1) * Q + (xc - 1)) * 3;
                float4* pob = ob + (y0 * P + xc) * 4 + q;
                const int* pv = p2v + y0 * P + xc;
                float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
                int v = pv[0];
                if (BWD) { if (v >= 0) e = ldg4(ext + (size_t)v * F); }
                int y = y0;
                while (y + 3 <= y1) {                                          // three rows per trip: the window rotates by name
                    DSB_TILE_ROW(a0, a1, a2, b0, b1, b2, c0, c1, c2)
                    DSB_TILE_ROW(b0, b1, b2, c0, c1, c2, a0, a1, a2)
                    DSB_TILE_ROW(c0, c1, c2, a0, a1, a2, b0, b1, b2)
                }
                if (y < y1) {
                    DSB_TILE_ROW(a0, a1, a2, b0, b1, b2, c0, c1, c2)
                    if (y < y1) DSB_TILE_ROW(b0, b1, b2, c0, c1, c2, a0, a1, a2)
                }
            }
            if (!last) __syncthreads();
        }
    }
}

size_t tile_smem(int P) {
    const int Q = P - 2;
    return (size_t)2 * P * P * 64 + (size_t)Q * Q * kWF * 4 + (size_t)P * P * 4;
}

}  // namespace
}  // namespace dsb200

using namespace dsb200;

extern "C" int dsb200_cheb_tile(const int32_t* tile_p2v, const float* tile_w, int ntiles, int T, int S, int B, int M, int F,
                                int backward, int monomial, const float* x, float* out, void* stream)
{
    DSB_REQUIRE(T == kT, "cheb_tile: the tile side is 16");
    DSB_REQUIRE(S >= 1 && S <= kMaxS, "cheb_tile: 1 <= K - 1 <= 7");
    DSB_REQUIRE(F > 0 && F % 16 == 0, "cheb_tile: the feature count must be a multiple of 16");
    DSB_REQUIRE(ntiles > 0 && B > 0 && (long long)ntiles * kT * kT == M, "cheb_tile: M must be ntiles * 256");
    DSB_REQUIRE(((((uintptr_t)x) | ((uintptr_t)out) | ((uintptr_t)tile_p2v) | ((uintptr_t)tile_w)) & 15) == 0,
                "cheb_tile: pointers must be 16-byte aligned");
    TileArgs a;
    a.p2v = tile_p2v; a.wts = tile_w; a.x = x; a.out = out;
    a.nt = ntiles; a.B = B; a.M = M; a.F = F; a.S = S; a.P = kT + 2 * S; a.NC = F / 16; a.mono = monomial ? 1 : 0;
    a.plane = (long long)B * M * F;
    a.nitems = (long long)ntiles * B * a.NC;
    const size_t smem = tile_smem(a.P);
    DSB_REQUIRE(smem <= 227 * 1024, "cheb_tile: the image does not fit in shared memory");
    const int ctas = DSB_ENV_INT("DSB200_TILE_CTAS", 2) == 1 ? 1 : (smem * 2 + 2048 <= 227 * 1024 ? 2 : 1);
    const long long cap = (long long)ctas * kNumSMs;
    const int grid = (int)(a.nitems < cap ? a.nitems : cap);
    cudaStream_t s = (cudaStream_t)stream;
#define DSB_TILE_LAUNCH(BW, SSV)                                                                                          \
    {                                                                                                                     \
        static size_t configured = 0;                                                                                     \
        if (smem > configured) {                                                                                          \
            cudaError_t e = cudaFuncSetAttribute(cheb_tile_kernel<BW, SSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return (int)e; }                                    \
            configured = smem;                                                                                            \
        }                                                                                                                 \
        cheb_tile_kernel<BW, SSV><<<grid, kThreads, smem, s>>>(a);                                                        \
    }
#define DSB_TILE_DISPATCH(BW)                                                                                             \
    switch (S) {                                                                                                          \
        case 1: DSB_TILE_LAUNCH(BW, 1) break;                                                                             \
        case 2: DSB_TILE_LAUNCH(BW, 2) break;                                                                             \
        case 3: DSB_TILE_LAUNCH(BW, 3) break;                                                                             \
        case 4: DSB_TILE_LAUNCH(BW, 4) break;                                                                             \
        default: DSB_TILE_LAUNCH(BW, 0) break;                                                                            \
    }
    if (backward) { DSB_TILE_DISPATCH(true) } else { DSB_TILE_DISPATCH(false) }
    DSB_LAUNCH_CHECK();
}
