// K2, general fp32 path for the shapes the tensor-core kernels do not take (feature counts that are not multiples of 8 -- the GHCN
// stack's 5 / 50 / 100 features -- and layers too wide for TMEM or resident weights -- the climate decoder's 4096 x 512 filters):
// one register-tiled FFMA GEMM core for the three products of cgcnn.chebyshev5 (reference models.py:1062-1078 and its autodiff),
//   FWD      Y[r, fo]        = sum_{k, fi} X_k[r, fi] W[fi*K + k, fo]
//   BWD_DATA dX_k[r, fi]     = sum_fo dY[r, fo] W[fi*K + k, fo]
//   WGRAD    dW[fi*K+k, fo] += sum_r X_k[r, fi] dY[r, fo]
// written as C[m, n] = sum_t A(m, t) B(t, n) with loaders that address the planar operands in place (the K orders are K separate
// [R, Fin] planes, never concatenated).  CTA tile 128 x (16 TN) x 16, 256 threads, 8 x TN outputs per thread; the next k-tile is
// fetched into registers while the current one is multiplied out of shared memory (one barrier per k-tile).  Exact fp32
// arithmetic (no TF32), any shape, vector loads where the shape allows.  Replaces the 128 x 64 / 64 x 64 single-buffered kernels
// of round 1 (20 % of the FFMA peak on the GHCN layers, profiles/r02_trace_step_cfg5.txt).
#include "common.cuh"

namespace dsb200 {

constexpr int GM = 128, GK = 16, GPAD = 4;
enum { G_FWD = 0, G_BWD = 1, G_WGRAD = 2 };

struct GemmArgs {
    const float* a0;        // FWD / WGRAD: plane 0 of X; BWD: dY
    const float* ak;        // FWD / WGRAD: planes 1..K-1 of X
    const float* b;         // FWD / BWD: W [Fin*K, Fout]; WGRAD: dY [R, Fout]
    float* c;               // FWD: Y [R, Fout]; BWD: G [K][R][Fin]; WGRAD: dW [Fin*K, Fout]
    double* bn_sums;        // FWD only, may be NULL
    long long R;
    int Fin, K, Fout;
    long long M;            // rows of C: R (FWD / BWD), K*Fin (WGRAD)
    int N;                  // columns of C: Fout (FWD / WGRAD), K*Fin (BWD)
    long long T;            // reduction length: K*Fin (FWD), Fout (BWD), R (WGRAD)
    long long t_per_split;  // WGRAD: rows of the reduction per blockIdx.z
};

__device__ __forceinline__ const float* plane_of(const GemmArgs& g, int k) {
    return k == 0 ? g.a0 : g.ak + (long long)(k - 1) * g.R * g.Fin;
}

// A(m, t) for 4 consecutive t (FWD, BWD: operands contiguous along t) or 4 consecutive m (WGRAD)
template <int MODE>
__device__ __forceinline__ void load_a4(const GemmArgs& g, long long m, long long t, long long tend, float (&v)[4])
{
    v[0] = v[1] = v[2] = v[3] = 0.f;
    if (MODE == G_FWD) {                         // m = row r, t = q = k*Fin + fi
        if (m >= g.M) return;
        const int q = (int)t;
        const int k = q / g.Fin, fi = q - k * g.Fin;
        if ((g.Fin & 3) == 0 && q + 3 < (int)tend) {                       // 4 features of one plane, 16-byte aligned
            const float4 w = ldg4(plane_of(g, k) + m * g.Fin + fi);
            v[0] = w.x; v[1] = w.y; v[2] = w.z; v[3] = w.w;
        } else {
            int kk = k, ff = fi;
            const float* p = plane_of(g, kk) + m * g.Fin + ff;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (q + j < (int)tend) v[j] = __ldg(p);
                ++p;
                if (++ff == g.Fin) { ff = 0; ++kk; p = plane_of(g, kk) + m * g.Fin; }
            }
        }
    } else if (MODE == G_BWD) {                  // m = row r, t = fo
        if (m >= g.M) return;
        const float* p = g.a0 + m * g.Fout + t;
        if ((g.Fout & 3) == 0 && t + 3 < tend) { const float4 w = ldg4(p); v[0] = w.x; v[1] = w.y; v[2] = w.z; v[3] = w.w; }
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (t + j < tend) v[j] = __ldg(p + j);
        }
    } else {                                     // WGRAD: m = q (4 consecutive), t = row r
        if (t >= tend) return;
        const int q = (int)m;
        const int k = q / g.Fin, fi = q - k * g.Fin;
        if ((g.Fin & 3) == 0 && q + 3 < (int)g.M) {
            const float4 w = ldg4(plane_of(g, k) + t * g.Fin + fi);
            v[0] = w.x; v[1] = w.y; v[2] = w.z; v[3] = w.w;
        } else {
            int kk = k, ff = fi;
            const float* p = plane_of(g, kk) + t * g.Fin + ff;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (q + j < (int)g.M) v[j] = __ldg(p);
                ++p;
                if (++ff == g.Fin) { ff = 0; ++kk; p = plane_of(g, kk) + t * g.Fin; }
            }
        }
    }
}

// B(t, n) for 4 consecutive n (FWD, WGRAD) or 4 consecutive t (BWD: W is contiguous along fo = t)
template <int MODE>
__device__ __forceinline__ void load_b4(const GemmArgs& g, long long t, long long tend, int n, float (&v)[4])
{
    v[0] = v[1] = v[2] = v[3] = 0.f;
    if (MODE == G_FWD) {                         // t = q = k*Fin + fi -> row fi*K + k of W
        if (t >= tend) return;
        const int q = (int)t;
        const int k = q / g.Fin, fi = q - k * g.Fin;
        const float* p = g.b + ((long long)fi * g.K + k) * g.Fout + n;
        if ((g.Fout & 3) == 0 && n + 3 < g.N) { const float4 w = ldg4(p); v[0] = w.x; v[1] = w.y; v[2] = w.z; v[3] = w.w; }
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (n + j < g.N) v[j] = __ldg(p + j);
        }
    } else if (MODE == G_BWD) {                  // n = k*Fin + fi (one column), t = fo (4 consecutive)
        if (n >= g.N) return;
        const int k = n / g.Fin, fi = n - k * g.Fin;
        const float* p = g.b + ((long long)fi * g.K + k) * g.Fout + t;
        if ((g.Fout & 3) == 0 && t + 3 < tend) { const float4 w = ldg4(p); v[0] = w.x; v[1] = w.y; v[2] = w.z; v[3] = w.w; }
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (t + j < tend) v[j] = __ldg(p + j);
        }
    } else {                                     // WGRAD: t = row r, n = fo
        if (t >= tend) return;
        const float* p = g.b + t * g.Fout + n;
        if ((g.Fout & 3) == 0 && n + 3 < g.N) { const float4 w = ldg4(p); v[0] = w.x; v[1] = w.y; v[2] = w.z; v[3] = w.w; }
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (n + j < g.N) v[j] = __ldg(p + j);
        }
    }
}

// TN = output columns per thread: the CTA tile is 128 x (16 * TN)
template <int MODE, int TN>
__global__ void __launch_bounds__(256, 2)
sgemm_kernel(const GemmArgs g)
{
    constexpr int GN = 16 * TN;
    constexpr int ALOADS = GM * GK / 1024;                 // float4 groups of the A tile per thread
    constexpr int BLOADS = (GK * GN + 1023) / 1024;        // float4 groups of the B tile per thread
    __shared__ __align__(16) float As[2][GK][GM + GPAD];
    __shared__ __align__(16) float Bs[2][GK][GN + GPAD];
    __shared__ float red[2][GN];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long m0 = (long long)blockIdx.x * GM;
    const int n0 = blockIdx.y * GN;
    long long tbeg = 0, tend = g.T;
    if (MODE == G_WGRAD) {
        tbeg = (long long)blockIdx.z * g.t_per_split;
        tend = tbeg + g.t_per_split < g.T ? tbeg + g.t_per_split : g.T;
    }

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // ---- tile loaders: global -> registers, registers -> shared memory (As[kk][m], Bs[kk][n])
    float ra[ALOADS][4], rb[BLOADS][4];
    auto fetch = [&](long long t0) {
#pragma unroll
        for (int l = 0; l < ALOADS; ++l) {
            if (MODE == G_WGRAD) load_a4<MODE>(g, m0 + (tid & 31) * 4, t0 + (tid >> 5) + 8 * l, tend, ra[l]);   // 4 consecutive q of one row
            else load_a4<MODE>(g, m0 + (tid >> 1), t0 + (tid & 1) * 4 + 8 * l, tend, ra[l]);                     // 4 consecutive t of one row
        }
#pragma unroll
        for (int l = 0; l < BLOADS; ++l) {
            const int e = tid + l * 256;
            if (MODE == G_BWD) { if (e * 4 < GK * GN) load_b4<MODE>(g, t0 + (e & 3) * 4, tend, n0 + (e >> 2), rb[l]); }
            else { if (e * 4 < GK * GN) load_b4<MODE>(g, t0 + e / (GN / 4), tend, n0 + (e % (GN / 4)) * 4, rb[l]); }
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int l = 0; l < ALOADS; ++l) {
            if (MODE == G_WGRAD) *reinterpret_cast<float4*>(&As[buf][(tid >> 5) + 8 * l][(tid & 31) * 4]) = make_float4(ra[l][0], ra[l][1], ra[l][2], ra[l][3]);
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) As[buf][(tid & 1) * 4 + 8 * l + j][tid >> 1] = ra[l][j];
            }
        }
#pragma unroll
        for (int l = 0; l < BLOADS; ++l) {
            const int e = tid + l * 256;
            if (e * 4 >= GK * GN) continue;
            if (MODE == G_BWD) {
#pragma unroll
                for (int j = 0; j < 4; ++j) Bs[buf][(e & 3) * 4 + j][e >> 2] = rb[l][j];
            } else *reinterpret_cast<float4*>(&Bs[buf][e / (GN / 4)][(e % (GN / 4)) * 4]) = make_float4(rb[l][0], rb[l][1], rb[l][2], rb[l][3]);
        }
    };

    int buf = 0;
    fetch(tbeg);
    stash(0);
    __syncthreads();
    for (long long t0 = tbeg; t0 < tend; t0 += GK) {
        const bool more = t0 + GK < tend;
        if (more) fetch(t0 + GK);                           // global loads in flight during the arithmetic below
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[TN];
            if (TN % 4 == 0) {
#pragma unroll
                for (int j = 0; j < TN; j += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * TN + j]);
                    b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j) b[j] = Bs[buf][kk][tx * TN + j];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) {
            stash(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }

    // ---- epilogue
    const int nb = n0 + tx * TN;
    if (MODE == G_FWD) {
        const bool vec = (g.Fout & 3) == 0 && (TN & 3) == 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long r = m0 + ty * 8 + i;
            if (r >= g.M) continue;
            float* cp = g.c + r * g.Fout + nb;
#pragma unroll
            for (int j = 0; j < TN; j += (TN % 4 == 0 ? 4 : 1)) {
                if (vec && nb + j + 3 < g.N) st4(cp + j, make_float4(acc[i][j], acc[i][j + 1 < TN ? j + 1 : j], acc[i][j + 2 < TN ? j + 2 : j], acc[i][j + 3 < TN ? j + 3 : j]));
                else {
#pragma unroll
                    for (int e = 0; e < (TN % 4 == 0 ? 4 : 1); ++e) if (nb + j + e < g.N) cp[j + e] = acc[i][j + e < TN ? j + e : j];
                }
            }
        }
        if (g.bn_sums) {                                       // uniform branch: sum_r Y and sum_r Y^2 per output feature
            if (tid < GN) { red[0][tid] = 0.f; red[1][tid] = 0.f; }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                float s = 0.f, ss = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (m0 + ty * 8 + i < g.M) { s += acc[i][j]; ss = fmaf(acc[i][j], acc[i][j], ss); }
                atomicAdd(&red[0][tx * TN + j], s);
                atomicAdd(&red[1][tx * TN + j], ss);
            }
            __syncthreads();
            if (tid < GN && n0 + tid < g.N) {
                atomicAdd(g.bn_sums + n0 + tid, (double)red[0][tid]);
                atomicAdd(g.bn_sums + g.N + n0 + tid, (double)red[1][tid]);
            }
        }
    } else if (MODE == G_BWD) {
        const long long plane = g.R * (long long)g.Fin;        // n = k*Fin + fi -> plane k, column fi of G
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long r = m0 + ty * 8 + i;
            if (r >= g.M) continue;
            int k = nb / g.Fin, fi = nb - k * g.Fin;
            float* cp = g.c + k * plane + r * g.Fin + fi;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                if (nb + j < g.N) *cp = acc[i][j];
                ++cp;
                if (++fi == g.Fin) { fi = 0; cp += plane - g.Fin; }
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long q = m0 + ty * 8 + i;
            if (q >= g.M) continue;
            const int k = (int)q / g.Fin, fi = (int)q - k * g.Fin;
            float* cp = g.c + ((long long)fi * g.K + k) * g.Fout + nb;
#pragma unroll
            for (int j = 0; j < TN; ++j) if (nb + j < g.N) atomicAdd(cp + j, acc[i][j]);
        }
    }
}

template <int MODE>
static int sgemm_launch(const GemmArgs& g, unsigned splits, cudaStream_t s)
{
    const unsigned gm = (unsigned)((g.M + GM - 1) / GM);
#define DSB_SGEMM(TNV) do { dim3 grid(gm, (unsigned)((g.N + 16 * TNV - 1) / (16 * TNV)), splits); \
                            sgemm_kernel<MODE, TNV><<<grid, 256, 0, s>>>(g); } while (0)
    if (g.N > 64) DSB_SGEMM(8);
    else if (g.N > 32) DSB_SGEMM(4);
    else if (g.N > 16) DSB_SGEMM(2);
    else DSB_SGEMM(1);
#undef DSB_SGEMM
    DSB_LAUNCH_CHECK();
}

int contract_fwd_ffma(const float* x0, const float* xk, const float* W, float* Y, long long R, int Fin, int K, int Fout,
                      double* bn_sums, cudaStream_t s)
{
    GemmArgs g;
    g.a0 = x0; g.ak = xk; g.b = W; g.c = Y; g.bn_sums = bn_sums; g.R = R; g.Fin = Fin; g.K = K; g.Fout = Fout;
    g.M = R; g.N = Fout; g.T = (long long)K * Fin; g.t_per_split = g.T;
    return sgemm_launch<G_FWD>(g, 1, s);
}

int contract_bwd_data_ffma(const float* dY, const float* W, float* G, long long R, int Fin, int K, int Fout, cudaStream_t s)
{
    GemmArgs g;
    g.a0 = dY; g.ak = nullptr; g.b = W; g.c = G; g.bn_sums = nullptr; g.R = R; g.Fin = Fin; g.K = K; g.Fout = Fout;
    g.M = R; g.N = K * Fin; g.T = Fout; g.t_per_split = g.T;
    return sgemm_launch<G_BWD>(g, 1, s);
}

int contract_bwd_weight_ffma(const float* x0, const float* xk, const float* dY, float* dW, long long R, int Fin, int K, int Fout,
                             cudaStream_t s)
{
    GemmArgs g;
    g.a0 = x0; g.ak = xk; g.b = dY; g.c = dW; g.bn_sums = nullptr; g.R = R; g.Fin = Fin; g.K = K; g.Fout = Fout;
    g.M = (long long)K * Fin; g.N = Fout; g.T = R;
    // split the rows so that ~4 CTAs per SM are in flight; partial tiles are merged with atomics
    const int tn = Fout > 64 ? 128 : Fout > 32 ? 64 : Fout > 16 ? 32 : 16;
    const long long tiles = ((g.M + GM - 1) / GM) * ((Fout + tn - 1) / tn);
    long long splits = (4LL * kNumSMs + tiles - 1) / tiles;
    const long long max_splits = (R + 8 * GK - 1) / (8 * GK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    long long per = (R + splits - 1) / splits;
    per = (per + GK - 1) / GK * GK;
    splits = (R + per - 1) / per;
    g.t_per_split = per;
    return sgemm_launch<G_WGRAD>(g, (unsigned)splits, s);
}

}  // namespace dsb200
