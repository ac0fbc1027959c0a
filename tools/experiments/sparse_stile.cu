// K1: staged row tile + tensor-core blocks.  out[b] = alpha * L x[b] + beta * y[b] + gamma * z[b]   (models.py:1056-1061)
// The two experiments that each solved half of the problem, combined:
//   * sparse_tiled.cu stages a row tile and its halo ring in shared memory with cp.async (memory-level parallelism
//     without registers) but then spends the L1 / shared data path on 9 gathers per output row;
//   * sparse_blk.cu turns 16 rows x their <= 8*KS distinct columns into dense mma.sync TF32 blocks (2.25x fewer operand
//     bytes, 3x fewer instructions) but fetches them from global memory at 16 warps per SM and is latency bound.
// Here a CTA (8 warps) owns a 128-row tile x NS samples x a 16-feature slab: the x rows of the tile and of its ring-1
// halo (and the y rows of the tile) are staged once, then warp w multiplies block w of the tile: B fragments come from
// shared memory, A fragments (hi / lo) stay in registers for all NS samples, 3xTF32 with the truncation split.
// Host side (graph.py: StagedBlockPlan): per tile the halo row list, per block the LOCAL column list and the A fragments.
#include "common.cuh"

namespace dsb200 {

constexpr int kStRows = 128;               // rows per tile = 8 blocks of 16
// Staged layout: per sample two planes (feature halves 0-7 / 8-15) of [rows][8 floats].  A B fragment reads 8 consecutive
// floats of 4 rows of one plane: conflict-free when the 4 local row indices differ modulo 4 (runs of consecutive columns).
constexpr int kStSlab = 16;                // features per CTA

struct StileArgs {
    const int4* tinfo;         // per tile: (first ext entry, n1 = tile + ring-1 rows, 0, 0)
    const int32_t* ext;        // global ids of the ring-1 rows of every tile
    const uint16_t* lcol;      // [nblocks][8*KS] local column of every block column
    const float4* afrag;       // [nblocks][KS][32] A fragments
    const float *x, *y, *z;
    float* out;
    int B, M, F, ntiles, nblocks, n1cap, NS, nslab;
    float alpha, beta, gamma;
};

__device__ __forceinline__ void st_cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void st_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int KS>
__global__ void __launch_bounds__(256, 2)
spmm_stile_kernel(const __grid_constant__ StileArgs a)
{
    extern __shared__ float sm[];
    float* Xs = sm;                                          // [NS][2][n1cap][8]
    float* Ys = sm + (size_t)a.NS * a.n1cap * 16;            // [NS][2][128][8]   (only when y != NULL)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int sgroups = (a.B + a.NS - 1) / a.NS;
    const long long items = (long long)a.ntiles * sgroups * a.nslab;
    const long long sample = (long long)a.M * a.F;
    for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        const int slab = (int)(it % a.nslab);
        const long long it2 = it / a.nslab;
        const int sg = (int)(it2 % sgroups), tile = (int)(it2 / sgroups);
        const int4 inf = __ldg(a.tinfo + tile);
        const int r0 = tile * kStRows;
        const int nt = min(kStRows, a.M - r0);
        const int n1 = inf.y;
        const int32_t* ext = a.ext + inf.x;
        const int b0 = sg * a.NS, nb = min(a.NS, a.B - b0);
        const int col0 = slab * kStSlab;
        __syncthreads();                                     // the previous item's readers are done with the buffers
        // ---- stage: 4 x 16-byte chunks per row; thread = (row mod 64, chunk), no divisions
        {
            const int c = threadIdx.x & 3, half = c >> 1, cc = (c & 1) * 4;
            for (int r = threadIdx.x >> 2; r < n1; r += 64) {
                const int gr = r < nt ? r0 + r : __ldg(ext + r - nt);
                const float* src = a.x + (long long)b0 * sample + (long long)gr * a.F + col0 + 4 * c;
                float* dst = Xs + ((size_t)half * a.n1cap + r) * 8 + cc;
                for (int s = 0; s < nb; ++s) st_cp_async16(dst + (size_t)s * a.n1cap * 16, src + s * sample);
            }
            if (a.y) {
                for (int r = threadIdx.x >> 2; r < nt; r += 64) {
                    const float* src = a.y + (long long)b0 * sample + (long long)(r0 + r) * a.F + col0 + 4 * c;
                    float* dst = Ys + ((size_t)half * kStRows + r) * 8 + cc;
                    for (int s = 0; s < nb; ++s) st_cp_async16(dst + (size_t)s * kStRows * 16, src + s * sample);
                }
            }
        }
        // ---- this warp's block of L (overlaps the copies)
        const int blk = tile * 8 + warp;
        const bool has = blk < a.nblocks && warp * 16 < nt;
        uint32_t ah[KS][4], al[KS][4];
        int off[KS][2];
        if (has) {
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                const float4 q = __ldg(a.afrag + ((long long)blk * KS + s) * 32 + lane);
                const float av[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    ah[s][i] = __float_as_uint(av[i]);                                   // read truncated by the tensor core
                    al[s][i] = __float_as_uint(av[i] - __uint_as_float(ah[s][i] & 0xffffe000u));
                }
                off[s][0] = (int)__ldg(a.lcol + (long long)blk * (8 * KS) + 8 * s + t) * 8 + g;
                off[s][1] = (int)__ldg(a.lcol + (long long)blk * (8 * KS) + 8 * s + t + 4) * 8 + g;
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
        if (!has) continue;
        const int lr0 = warp * 16 + g, lr1 = lr0 + 8;                                    // local rows of this lane's results
        const bool ok0 = lr0 < nt, ok1 = lr1 < nt;
        for (int s = 0; s < nb; ++s) {
            const float* xs = Xs + (size_t)s * a.n1cap * 16;
            const int plane = a.n1cap * 8;
            float c[2][3][4];
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int j = 0; j < 3; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) c[n][j][i] = 0.f;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
                for (int n = 0; n < 2; ++n) {
                    const float v0 = xs[off[ks][0] + n * plane], v1 = xs[off[ks][1] + n * plane];
                    const uint32_t h0 = __float_as_uint(v0), h1 = __float_as_uint(v1);
                    const uint32_t l0 = __float_as_uint(v0 - __uint_as_float(h0 & 0xffffe000u));
                    const uint32_t l1 = __float_as_uint(v1 - __uint_as_float(h1 & 0xffffe000u));
                    st_mma(c[n][0], ah[ks], h0, h1);
                    st_mma(c[n][1], ah[ks], l0, l1);
                    st_mma(c[n][2], al[ks], h0, h1);
                }
            }
            const long long gb = (b0 + s) * sample + col0;
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                float r[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) r[i] = a.alpha * ((c[n][1][i] + c[n][2][i]) + c[n][0][i]);
                const int cc = 8 * n + 2 * t;
                if (a.y) {
                    const float* ys = Ys + (size_t)s * kStRows * 16 + (size_t)n * kStRows * 8 + 2 * t;
                    if (ok0) { const float2 w = *reinterpret_cast<const float2*>(ys + lr0 * 8); r[0] = fmaf(a.beta, w.x, r[0]); r[1] = fmaf(a.beta, w.y, r[1]); }
                    if (ok1) { const float2 w = *reinterpret_cast<const float2*>(ys + lr1 * 8); r[2] = fmaf(a.beta, w.x, r[2]); r[3] = fmaf(a.beta, w.y, r[3]); }
                }
                const long long o0 = gb + (long long)(r0 + lr0) * a.F + cc, o1 = o0 + 8LL * a.F;
                if (a.z) {
                    if (ok0) { const float2 w = *reinterpret_cast<const float2*>(a.z + o0); r[0] = fmaf(a.gamma, w.x, r[0]); r[1] = fmaf(a.gamma, w.y, r[1]); }
                    if (ok1) { const float2 w = *reinterpret_cast<const float2*>(a.z + o1); r[2] = fmaf(a.gamma, w.x, r[2]); r[3] = fmaf(a.gamma, w.y, r[3]); }
                }
                if (ok0) *reinterpret_cast<float2*>(a.out + o0) = make_float2(r[0], r[1]);
                if (ok1) *reinterpret_cast<float2*>(a.out + o1) = make_float2(r[2], r[3]);
            }
        }
    }
}

template <int KS>
static int launch_stile(StileArgs& a, cudaStream_t s)
{
    int NS = 4;
    if (const char* e = getenv("DSB200_ST_NS")) NS = atoi(e) > 0 ? atoi(e) : 4;
    if (NS > a.B) NS = a.B;
    a.NS = NS;
    const size_t smem = ((size_t)NS * a.n1cap + (a.y ? (size_t)NS * kStRows : 0)) * 16 * sizeof(float);
    DSB_REQUIRE(smem <= 112 * 1024, "spmm_stile: tile + halo does not fit shared memory");
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(spmm_stile_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    const long long items = (long long)a.ntiles * ((a.B + NS - 1) / NS) * a.nslab;
    const long long cap = 2LL * kNumSMs;
    const int grid = (int)(items < cap ? items : cap);
    spmm_stile_kernel<KS><<<grid, 256, smem, s>>>(a);
    DSB_LAUNCH_CHECK();
}

}  // namespace dsb200

using namespace dsb200;

extern "C" int dsb200_spmm_stile(const int32_t* tile_info, const int32_t* tile_ext, const uint16_t* blk_lcols, const float* blk_frags,
                                 int ntiles, int nblocks, int ksteps, int n1cap, int B, int M, int F,
                                 const float* x, const float* y, const float* z, float* out,
                                 float alpha, float beta, float gamma, void* stream)
{
    DSB_REQUIRE(tile_info && tile_ext && blk_lcols && blk_frags && x && out && B > 0 && M > 0, "spmm_stile: null pointer");
    DSB_REQUIRE(ntiles == (M + kStRows - 1) / kStRows && nblocks == (M + 15) / 16, "spmm_stile: the plan does not match M");
    DSB_REQUIRE(F > 0 && F % kStSlab == 0, "spmm_stile: F must be a multiple of 16");
    DSB_REQUIRE(x != out, "spmm_stile: out must not alias x (halo rows are read by other tiles)");
    const uintptr_t al = (uintptr_t)x | (uintptr_t)y | (uintptr_t)z | (uintptr_t)out | (uintptr_t)blk_frags;
    DSB_REQUIRE((al & 15) == 0, "spmm_stile: tensors must be 16-byte aligned");
    if (beta == 0.f) y = nullptr;
    if (gamma == 0.f) z = nullptr;
    StileArgs a;
    a.tinfo = (const int4*)tile_info; a.ext = tile_ext; a.lcol = blk_lcols; a.afrag = (const float4*)blk_frags;
    a.x = x; a.y = y; a.z = z; a.out = out;
    a.B = B; a.M = M; a.F = F; a.ntiles = ntiles; a.nblocks = nblocks; a.n1cap = n1cap; a.nslab = F / kStSlab;
    a.alpha = alpha; a.beta = beta; a.gamma = gamma;
    cudaStream_t s = (cudaStream_t)stream;
    if (ksteps == 5) return launch_stile<5>(a, s);
    if (ksteps == 6) return launch_stile<6>(a, s);
    if (ksteps == 7) return launch_stile<7>(a, s);
    DSB_REQUIRE(false, "spmm_stile: 5, 6 or 7 k-steps");
}
