// K1 on the tensor cores: block formulation of the sparse-Laplacian product
//   out[b] = alpha * L x[b] + beta * y[b] + gamma * z[b]        (same contract as dsb200_spmm_ell; models.py:1056-1061)
// The gather kernel (sparse.cu) is bound by L1 requests: 9 scattered row gathers per output row.  Here 16 consecutive
// rows (a 4 x 4 patch of the sphere in HEALPix nested order) are ONE small dense product: their at most 8*KS distinct
// neighbour columns (36 for an interior patch) form a [16 x 8*KS] block A of L, and
//        out[16 rows, F] = A . x[neighbour rows, F]
// runs as mma.sync.m16n8k8 TF32 MMAs with the 3xTF32 split (A_hi B_hi + A_hi B_lo + A_lo B_hi, fp32 accumulate), which
// keeps fp32-level accuracy.  Every neighbour row is fetched ONCE per block instead of once per referencing row (2.25x
// fewer bytes through L1, and as 32-byte segments), and the A fragments of a block are reused for all samples.
// Host side (graph.py: BlockPlan): per block the column list u[8*KS] (global row ids; padding = first row of the block
// with zero weights) and A in MMA fragment order (KS x 32 lanes x float4).
#include "common.cuh"

namespace dsb200 {

__device__ __forceinline__ uint32_t f2tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// one warp = one 16-row block x `spw` consecutive samples; 8 warps per CTA take 8 consecutive blocks (a 128-row tile)
template <int KS, int NP>
__global__ void __launch_bounds__(256)
spmm_blk_kernel(const int32_t* __restrict__ ucol, const float4* __restrict__ afrag, int nblocks, int B, int M, int F, int spw,
                const float* __restrict__ x, const float* y, const float* z, float* out, float alpha, float beta, float gamma)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int sgroups = (B + spw - 1) / spw;
    const long long items = (long long)(nblocks + 7) / 8 * sgroups;           // (8-block tile, sample group)
    const int NT = F >> 3;
    const long long sample = (long long)M * F;
    for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        const int sg = (int)(it % sgroups);
        const int blk = (int)(it / sgroups) * 8 + warp;
        if (blk >= nblocks) continue;
        // ---- the block of L: fragments (hi / lo) and the row offsets this lane needs
        uint32_t ah[KS][4], al[KS][4];
        int off[KS][2];
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            const float4 a = __ldg(afrag + ((long long)blk * KS + s) * 32 + lane);
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ah[s][i] = f2tf32(av[i]);
                al[s][i] = f2tf32(av[i] - __uint_as_float(ah[s][i]));
            }
            off[s][0] = __ldg(ucol + (long long)blk * (8 * KS) + 8 * s + t) * F + g;
            off[s][1] = __ldg(ucol + (long long)blk * (8 * KS) + 8 * s + t + 4) * F + g;
        }
        const int r0 = blk * 16 + g, r1 = r0 + 8;
        const bool ok0 = r0 < M, ok1 = r1 < M;
        const int b_end = min(B, (sg + 1) * spw);
        for (int b = sg * spw; b < b_end; ++b) {
            const float* xb = x + b * sample;
            if (b + 1 < b_end) {
                // The next sample's neighbour rows (and its y rows) are requested now, without destination registers:
                // the kernel runs at 16 warps per SM, so HBM latency has to be hidden by distance, not by occupancy.
                const float* xn = xb + sample;
#pragma unroll
                for (int s = 0; s < KS; ++s) {
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(xn + off[s][0] - g + 8 * (g & 1)));
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(xn + off[s][1] - g + 8 * (g & 1)));
                }
                if (y && t == 0) {
                    const float* yn = y + (b + 1) * sample;
                    if (ok0) asm volatile("prefetch.global.L1 [%0];" :: "l"(yn + (long long)r0 * F));
                    if (ok1) asm volatile("prefetch.global.L1 [%0];" :: "l"(yn + (long long)r1 * F));
                }
            }
            for (int n = 0; n < NT; n += NP) {
                // NP column tiles at once: 2 * KS * NP loads in flight, 3 * NP independent accumulator chains
                float bv[NP][KS][2];
#pragma unroll
                for (int q = 0; q < NP; ++q) {
#pragma unroll
                    for (int s = 0; s < KS; ++s) {
                        bv[q][s][0] = __ldg(xb + off[s][0] + 8 * (n + q));
                        bv[q][s][1] = __ldg(xb + off[s][1] + 8 * (n + q));
                    }
                }
                float chh[NP][4], chl[NP][4], clh[NP][4];
#pragma unroll
                for (int q = 0; q < NP; ++q) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) { chh[q][i] = 0.f; chl[q][i] = 0.f; clh[q][i] = 0.f; }
                }
#pragma unroll
                for (int s = 0; s < KS; ++s) {
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        // cvt.rna.tf32 is ~10 SASS instructions; the tensor core ignores the low 13 mantissa bits of a TF32
                        // operand, so hi = the raw value (read truncated) and lo = the bits it drops, exact in fp32
                        const uint32_t h0 = __float_as_uint(bv[q][s][0]), h1 = __float_as_uint(bv[q][s][1]);
                        const uint32_t l0 = __float_as_uint(bv[q][s][0] - __uint_as_float(h0 & 0xffffe000u));
                        const uint32_t l1 = __float_as_uint(bv[q][s][1] - __uint_as_float(h1 & 0xffffe000u));
                        mma_tf32_16x8x8(chh[q], ah[s], h0, h1);
                        mma_tf32_16x8x8(chl[q], ah[s], l0, l1);
                        mma_tf32_16x8x8(clh[q], al[s], h0, h1);
                    }
                }
                // ---- epilogue: this lane holds (r0, 8n+2t..+1) and (r1, 8n+2t..+1) of every column tile
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    float c[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) c[i] = (chl[q][i] + clh[q][i]) + chh[q][i];
                    const long long o0 = b * sample + (long long)r0 * F + 8 * (n + q) + 2 * t;
                    const long long o1 = o0 + 8LL * F;
                    float2 v0 = make_float2(alpha * c[0], alpha * c[1]), v1 = make_float2(alpha * c[2], alpha * c[3]);
                    if (y) {
                        if (ok0) { const float2 w = *reinterpret_cast<const float2*>(y + o0); v0.x = fmaf(beta, w.x, v0.x); v0.y = fmaf(beta, w.y, v0.y); }
                        if (ok1) { const float2 w = *reinterpret_cast<const float2*>(y + o1); v1.x = fmaf(beta, w.x, v1.x); v1.y = fmaf(beta, w.y, v1.y); }
                    }
                    if (z) {
                        if (ok0) { const float2 w = *reinterpret_cast<const float2*>(z + o0); v0.x = fmaf(gamma, w.x, v0.x); v0.y = fmaf(gamma, w.y, v0.y); }
                        if (ok1) { const float2 w = *reinterpret_cast<const float2*>(z + o1); v1.x = fmaf(gamma, w.x, v1.x); v1.y = fmaf(gamma, w.y, v1.y); }
                    }
                    if (ok0) *reinterpret_cast<float2*>(out + o0) = v0;
                    if (ok1) *reinterpret_cast<float2*>(out + o1) = v1;
                }
            }
        }
    }
}

template <int KS>
static int launch_blk(const int32_t* ucol, const float* afrag, int nblocks, int B, int M, int F, const float* x, const float* y,
                      const float* z, float* out, float alpha, float beta, float gamma, cudaStream_t s)
{
    // samples per warp: reuse the A fragments as long as the machine stays full
    const long long tiles = (nblocks + 7) / 8;
    int spw = B;
    if (const char* e = getenv("DSB200_BLK_SPW")) spw = atoi(e) > 0 ? atoi(e) : B;
    else while (spw > 1 && tiles * ((B + spw - 1) / spw) < 4LL * kNumSMs) spw = (spw + 1) / 2;
    const long long items = tiles * ((B + spw - 1) / spw);
    int ctas = 2;
    if (const char* e = getenv("DSB200_BLK_CTAS")) ctas = atoi(e) > 0 ? atoi(e) : 2;
    const long long cap = (long long)kNumSMs * ctas;
    const int grid = (int)(items < cap ? items : cap);
    if ((F >> 3) % 2 == 0)
        spmm_blk_kernel<KS, 2><<<grid, 256, 0, s>>>(ucol, (const float4*)afrag, nblocks, B, M, F, spw, x, y, z, out, alpha, beta, gamma);
    else
        spmm_blk_kernel<KS, 1><<<grid, 256, 0, s>>>(ucol, (const float4*)afrag, nblocks, B, M, F, spw, x, y, z, out, alpha, beta, gamma);
    DSB_LAUNCH_CHECK();
}

}  // namespace dsb200

using namespace dsb200;

extern "C" int dsb200_spmm_blk(const int32_t* blk_cols, const float* blk_frags, int nblocks, int ksteps, int B, int M, int F,
                               const float* x, const float* y, const float* z, float* out,
                               float alpha, float beta, float gamma, void* stream)
{
    DSB_REQUIRE(blk_cols && blk_frags && x && out && B > 0 && M > 0, "spmm_blk: null pointer");
    DSB_REQUIRE(nblocks == (M + 15) / 16, "spmm_blk: the block plan does not match M");
    DSB_REQUIRE(F > 0 && F % 8 == 0, "spmm_blk: F must be a multiple of 8");
    DSB_REQUIRE(x != out, "spmm_blk: out must not alias x");
    DSB_REQUIRE((long long)M * F < 0x7fffffffLL, "spmm_blk: a sample must have fewer than 2^31 elements");
    const uintptr_t al = (uintptr_t)x | (uintptr_t)y | (uintptr_t)z | (uintptr_t)out;
    DSB_REQUIRE((al & 7) == 0 && (((uintptr_t)blk_frags) & 15) == 0, "spmm_blk: misaligned tensors");
    if (beta == 0.f) y = nullptr;
    if (gamma == 0.f) z = nullptr;
    cudaStream_t s = (cudaStream_t)stream;
    if (ksteps == 5) return launch_blk<5>(blk_cols, blk_frags, nblocks, B, M, F, x, y, z, out, alpha, beta, gamma, s);
    if (ksteps == 6) return launch_blk<6>(blk_cols, blk_frags, nblocks, B, M, F, x, y, z, out, alpha, beta, gamma, s);
    if (ksteps == 7) return launch_blk<7>(blk_cols, blk_frags, nblocks, B, M, F, x, y, z, out, alpha, beta, gamma, s);
    DSB_REQUIRE(false, "spmm_blk: 5, 6 or 7 k-steps (at most 56 distinct columns per 16-row block)");
}
