// K2 forward for NARROW layers (Fin = 16, Fout = 32: cosmo layer 2, the largest contraction of the step) without any
// shared-memory staging of the activations:  Y[r, fo] = sum_{k, fi} X_k[r, fi] W[fi*K + k, fo]   (models.py:1062-1078).
//
// Why: the tcgen05 row-GEMM (contract_tc.cu) moves every activation byte four times through shared memory (TMA write,
// split read, split write of the low part, two tensor-core reads) and is bound by the SM's load/store pipe at 1.8x its
// HBM time (tools/exp_gemm_parts.py).  The reduction depth per plane is only 16, so the legacy warp-level MMA
// (mma.sync.m16n8k8 TF32) is fast enough, and it takes its A operand from REGISTERS: a warp reads its 32 rows of a plane
// with four coalesced 16-byte loads per lane straight from global memory, splits hi / lo in registers (the tensor core
// reads a TF32 operand truncated, so hi is the value itself and lo = x - trunc(x)), and only the 10 KB weight image lives
// in shared memory, pre-arranged in fragment order (one conflict-free 16-byte read per k-step and column tile, shared by the
// two row tiles of the warp).  3xTF32: A_hi W_hi + A_lo W_hi + A_hi W_lo, fp32 accumulate.
//
// Fragment bookkeeping (PTX ISA, mma.m16n8k8 .tf32): lane = 4 g + t.  A: a0 = (row g, k t), a1 = (row g+8, k t), a2 = (row g,
// k t+4), a3 = (row g+8, k t+4);  B: b0 = (k t, n g), b1 = (k t+4, n g);  C: c0, c1 = (row g, n 2t, 2t+1), c2, c3 = row g+8.
// The k index of an MMA is a free permutation as long as A and B agree, and so is n between B and the output: a lane loads
// features 4t..4t+3 of its rows (16 bytes), feeds (4t, 4t+1) to the first k-step as its (t, t+4) and (4t+2, 4t+3) to the
// second, and column tile j of a pair (p, q) = (j / 2, j % 2) maps MMA column 2t' + e to output column 16p + 4t' + 2q + e,
// so that a lane ends up with 4 consecutive output columns per row and 16-column group: 16-byte stores, 64 contiguous
// bytes per row and instruction.  Batch-norm sums: per-lane partial column sums over all its rows, reduced once per CTA.
#include "common.cuh"
#include <stdlib.h>

namespace dsb200 {

constexpr int kMmaFin = 16, kMmaFout = 32, kMmaMaxK = 8;

__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256, 2)
rowgemm_mma_kernel(const float* __restrict__ x0, const float* __restrict__ xk, const float* __restrict__ W,
                   float* __restrict__ Y, double* __restrict__ bn_sums, long long R, int K)
{
    __shared__ float4 sW[kMmaMaxK * 2 * 4 * 32];               // [plane][k-step][column tile][lane] = (b0 hi, b1 hi, b0 lo, b1 lo)
    __shared__ float red[2 * kMmaFout];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = lane & 3, g = lane >> 2;
    for (int i = threadIdx.x; i < K * 2 * 4 * 32; i += 256) {
        const int l = i & 31, j = (i >> 5) & 3, s = (i >> 7) & 1, k = i >> 8;
        const int tt = l & 3, gg = l >> 2;
        const int fi0 = 4 * tt + 2 * s;                          // features this lane feeds to k indices (tt, tt + 4)
        const int col = 16 * (j >> 1) + 4 * (gg >> 1) + 2 * (j & 1) + (gg & 1);
        const float w0 = __ldg(W + (fi0 * K + k) * kMmaFout + col), w1 = __ldg(W + ((fi0 + 1) * K + k) * kMmaFout + col);
        uint32_t h0, h1;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h0) : "f"(w0));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h1) : "f"(w1));
        sW[i] = make_float4(__uint_as_float(h0), __uint_as_float(h1), w0 - __uint_as_float(h0), w1 - __uint_as_float(h1));
    }
    if (threadIdx.x < 2 * kMmaFout) red[threadIdx.x] = 0.f;
    __syncthreads();

    float s1[8], s2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
    const long long plane = R * kMmaFin;
    const long long nblk = (R + 31) / 32;
    for (long long blk = (long long)blockIdx.x * 8 + warp; blk < nblk; blk += (long long)gridDim.x * 8) {
        const long long row0 = blk * 32;
        float acc[2][4][4];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[m][j][e] = 0.f;
        // rows of this lane: row0 + 16 m + 8 h + g
        float4 v[2][2], vn[2][2];
        auto load = [&](int k, float4 (&dst)[2][2]) {
            const float* xp = (k == 0 ? x0 : xk + (long long)(k - 1) * plane) + 4 * t;
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const long long r = row0 + 16 * m + 8 * h + g;
                    dst[m][h] = r < R ? __ldg(reinterpret_cast<const float4*>(xp + r * kMmaFin)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
        };
        load(0, v);
        for (int k = 0; k < K; ++k) {
            if (k + 1 < K) load(k + 1, vn);
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                uint32_t ah[2][4], al[2][4];
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const float f0 = s ? v[m][0].z : v[m][0].x, f1 = s ? v[m][1].z : v[m][1].x;     // k index t:     rows g, g + 8
                    const float f2 = s ? v[m][0].w : v[m][0].y, f3 = s ? v[m][1].w : v[m][1].y;     // k index t + 4
                    const float f[4] = {f0, f1, f2, f3};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const uint32_t hb = __float_as_uint(f[e]) & 0xffffe000u;
                        ah[m][e] = hb;
                        al[m][e] = __float_as_uint(f[e] - __uint_as_float(hb));
                    }
                }
                const float4* wp = sW + ((k * 2 + s) * 4) * 32 + lane;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 wf = wp[j * 32];
                    const uint32_t bh0 = __float_as_uint(wf.x), bh1 = __float_as_uint(wf.y);
                    const uint32_t bl0 = __float_as_uint(wf.z), bl1 = __float_as_uint(wf.w);
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        mma_tf32_16x8x8(acc[m][j], al[m], bh0, bh1);         // small terms first
                        mma_tf32_16x8x8(acc[m][j], ah[m], bl0, bl1);
                        mma_tf32_16x8x8(acc[m][j], ah[m], bh0, bh1);
                    }
                }
            }
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int h = 0; h < 2; ++h) v[m][h] = vn[m][h];
        }
        // a lane owns columns 16 p + 4 t .. + 3 (p = 0, 1) of rows g, g + 8 of each row tile
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const long long r = row0 + 16 * m + 8 * h + g;
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    const float4 o = make_float4(acc[m][2 * p][2 * h], acc[m][2 * p][2 * h + 1], acc[m][2 * p + 1][2 * h], acc[m][2 * p + 1][2 * h + 1]);
                    if (r < R) *reinterpret_cast<float4*>(Y + r * kMmaFout + 16 * p + 4 * t) = o;
                    if (bn_sums) {                                           // rows past R are zero
                        s1[4 * p] += o.x; s1[4 * p + 1] += o.y; s1[4 * p + 2] += o.z; s1[4 * p + 3] += o.w;
                        s2[4 * p] = fmaf(o.x, o.x, s2[4 * p]); s2[4 * p + 1] = fmaf(o.y, o.y, s2[4 * p + 1]);
                        s2[4 * p + 2] = fmaf(o.z, o.z, s2[4 * p + 2]); s2[4 * p + 3] = fmaf(o.w, o.w, s2[4 * p + 3]);
                    }
                }
            }
    }
    if (bn_sums) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float a = s1[i], b = s2[i];
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
            if (g == 0) {
                const int col = 16 * (i >> 2) + 4 * t + (i & 3);
                atomicAdd(red + col, a);
                atomicAdd(red + kMmaFout + col, b);
            }
        }
        __syncthreads();
        if (threadIdx.x < 2 * kMmaFout) atomicAdd(bn_sums + threadIdx.x, (double)red[threadIdx.x]);
    }
}

// ---- the same computation with the activations arriving by 1-D bulk copies (TMA) through a shared-memory ring, so that many
// more bytes are in flight than a lane's registers can hold (the register version is latency bound: 4 loads per lane ahead).
// One producer warp; a pipeline item = (block of 256 consecutive rows, plane k) = 16 KB contiguous in memory; the 8 consumer
// warps own 32 rows each and read their A fragments from the stage with the same 16-byte pattern (8 rows x 64 bytes per
// quarter warp: conflict-free).
__device__ __forceinline__ uint32_t mp_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mp_bar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mp_saddr(bar)), "r"(count));
}
__device__ __forceinline__ void mp_bar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mp_saddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mp_bar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(mp_saddr(bar)) : "memory");
}
__device__ __forceinline__ void mp_bar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n .reg .pred p;\n W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D;\n bra W;\n D:\n}"
                 :: "r"(mp_saddr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mp_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(mp_saddr(dst)), "l"(src), "r"(bytes), "r"(mp_saddr(bar)) : "memory");
}

constexpr int kMpThreads = 9 * 32;

// MTW = row tiles of 16 per warp (2: 32 rows, two CTAs per SM; 4: 64 rows, one CTA per SM -- the weight fragments are then read
// half as often per row), kMpStages stages of 8 * 16 * MTW rows.
template <bool SUMS, int MTW, int kMpStages, bool POOL>
__device__ __forceinline__ void
rowgemm_mma_pipe_body(const float* __restrict__ x0, const float* __restrict__ xk, const float* __restrict__ W,
                      float* __restrict__ Y, double* __restrict__ bn_sums, long long R, int K, float* __restrict__ Ypool)
{
    constexpr int kMpRows = 8 * 16 * MTW, kMpStageBytes = kMpRows * kMmaFin * 4;
    extern __shared__ __align__(128) unsigned char smem[];
    float4* sW = reinterpret_cast<float4*>(smem + kMpStages * kMpStageBytes);           // [K][2][4][32]
    float* red = reinterpret_cast<float*>(sW + kMmaMaxK * 2 * 4 * 32);
    uint64_t* full = reinterpret_cast<uint64_t*>(red + 2 * kMmaFout);
    uint64_t* empty = full + kMpStages;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = lane & 3, g = lane >> 2;
    for (int i = threadIdx.x; i < K * 2 * 4 * 32; i += kMpThreads) {
        const int l = i & 31, j = (i >> 5) & 3, s = (i >> 7) & 1, k = i >> 8;
        const int tt = l & 3, gg = l >> 2;
        const int fi0 = 4 * tt + 2 * s;
        const int col = 16 * (j >> 1) + 4 * (gg >> 1) + 2 * (j & 1) + (gg & 1);
        const float w0 = __ldg(W + (fi0 * K + k) * kMmaFout + col), w1 = __ldg(W + ((fi0 + 1) * K + k) * kMmaFout + col);
        uint32_t h0, h1;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h0) : "f"(w0));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h1) : "f"(w1));
        sW[i] = make_float4(__uint_as_float(h0), __uint_as_float(h1), w0 - __uint_as_float(h0), w1 - __uint_as_float(h1));
    }
    if (threadIdx.x < 2 * kMmaFout) red[threadIdx.x] = 0.f;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kMpStages; ++s) { mp_bar_init(full + s, 1); mp_bar_init(empty + s, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long plane = R * kMmaFin;
    const long long nblk = (R + kMpRows - 1) / kMpRows;
    const long long b0 = nblk * blockIdx.x / gridDim.x, b1 = nblk * (blockIdx.x + 1) / gridDim.x;   // contiguous row blocks per CTA
    const int nitems = (int)(b1 - b0) * K;

    if (warp == 8) {
        if (lane == 0) {
            for (int it = 0; it < nitems; ++it) {
                const int s = it % kMpStages, u = it / kMpStages;
                const long long blk = b0 + it / K;
                const int k = it % K;
                if (u > 0) mp_bar_wait(empty + s, (uint32_t)(u - 1) & 1u);
                const long long row0 = blk * kMpRows;
                const long long rows = R - row0 < kMpRows ? R - row0 : kMpRows;
                const float* src = (k == 0 ? x0 : xk + (long long)(k - 1) * plane) + row0 * kMmaFin;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mp_bar_expect_tx(full + s, (uint32_t)rows * kMmaFin * 4);
                mp_bulk_g2s(smem + (size_t)s * kMpStageBytes, src, (uint32_t)rows * kMmaFin * 4, full + s);
            }
        }
        return;                                   // (the final reduction below uses a named barrier of the 8 consumer warps)
    }

    float s1[8], s2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
    int it = 0;
    for (long long blk = b0; blk < b1; ++blk) {
        const long long row0 = blk * kMpRows + warp * (16 * MTW);
        float acc[MTW][4][4];
#pragma unroll
        for (int m = 0; m < MTW; ++m)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[m][j][e] = 0.f;
        for (int k = 0; k < K; ++k, ++it) {
            const int s = it % kMpStages, u = it / kMpStages;
            mp_bar_wait(full + s, (uint32_t)u & 1u);
            const float4* A = reinterpret_cast<const float4*>(smem + (size_t)s * kMpStageBytes) + (warp * (16 * MTW)) * 4 + t;
            float4 v[MTW][2];
#pragma unroll
            for (int m = 0; m < MTW; ++m)
#pragma unroll
                for (int h = 0; h < 2; ++h)                               // rows past R: zeros (the stage holds stale data there)
                    v[m][h] = (row0 + 16 * m + 8 * h + g < R) ? A[(16 * m + 8 * h + g) * 4] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int s2i = 0; s2i < 2; ++s2i) {
                uint32_t ah[MTW][4], al[MTW][4];
#pragma unroll
                for (int m = 0; m < MTW; ++m) {
                    const float f0 = s2i ? v[m][0].z : v[m][0].x, f1 = s2i ? v[m][1].z : v[m][1].x;
                    const float f2 = s2i ? v[m][0].w : v[m][0].y, f3 = s2i ? v[m][1].w : v[m][1].y;
                    const float f[4] = {f0, f1, f2, f3};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const uint32_t hb = __float_as_uint(f[e]) & 0xffffe000u;
                        ah[m][e] = hb;
                        al[m][e] = __float_as_uint(f[e] - __uint_as_float(hb));
                    }
                }
                const float4* wp = sW + ((k * 2 + s2i) * 4) * 32 + lane;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 wf = wp[j * 32];
                    const uint32_t bh0 = __float_as_uint(wf.x), bh1 = __float_as_uint(wf.y);
                    const uint32_t bl0 = __float_as_uint(wf.z), bl1 = __float_as_uint(wf.w);
#pragma unroll
                    for (int m = 0; m < MTW; ++m) {
                        mma_tf32_16x8x8(acc[m][j], al[m], bh0, bh1);
                        mma_tf32_16x8x8(acc[m][j], ah[m], bl0, bl1);
                        mma_tf32_16x8x8(acc[m][j], ah[m], bh0, bh1);
                    }
                }
            }
            __syncwarp();                                                // every lane has consumed its registers of the stage
            if (lane == 0) mp_bar_arrive(empty + s);
        }
#pragma unroll
        for (int m = 0; m < MTW; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const long long r = row0 + 16 * m + 8 * h + g;
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    const float4 o = make_float4(acc[m][2 * p][2 * h], acc[m][2 * p][2 * h + 1], acc[m][2 * p + 1][2 * h], acc[m][2 * p + 1][2 * h + 1]);
                    if ((!POOL || Y != nullptr) && r < R) *reinterpret_cast<float4*>(Y + r * kMmaFout + 16 * p + 4 * t) = o;
                    if (POOL) {
                        // max-pool over 4 consecutive rows (g, g^1, g^2, g^3 = lanes 4 and 8 apart), on the raw y: it commutes with the
                        // batch norm / bias / activation that follow (non-decreasing per feature, models.py:1235-1243)
                        float4 q = o;
                        q.x = fmaxf(q.x, __shfl_xor_sync(0xffffffffu, q.x, 4)); q.y = fmaxf(q.y, __shfl_xor_sync(0xffffffffu, q.y, 4));
                        q.z = fmaxf(q.z, __shfl_xor_sync(0xffffffffu, q.z, 4)); q.w = fmaxf(q.w, __shfl_xor_sync(0xffffffffu, q.w, 4));
                        q.x = fmaxf(q.x, __shfl_xor_sync(0xffffffffu, q.x, 8)); q.y = fmaxf(q.y, __shfl_xor_sync(0xffffffffu, q.y, 8));
                        q.z = fmaxf(q.z, __shfl_xor_sync(0xffffffffu, q.z, 8)); q.w = fmaxf(q.w, __shfl_xor_sync(0xffffffffu, q.w, 8));
                        if ((g & 3) == 0 && r < R) *reinterpret_cast<float4*>(Ypool + (r >> 2) * kMmaFout + 16 * p + 4 * t) = q;
                    }
                    if (!SUMS) continue;
                    // rows past R are exact zeros: the sums need no predicate
                    s1[4 * p] += o.x; s1[4 * p + 1] += o.y; s1[4 * p + 2] += o.z; s1[4 * p + 3] += o.w;
                    s2[4 * p] = fmaf(o.x, o.x, s2[4 * p]); s2[4 * p + 1] = fmaf(o.y, o.y, s2[4 * p + 1]);
                    s2[4 * p + 2] = fmaf(o.z, o.z, s2[4 * p + 2]); s2[4 * p + 3] = fmaf(o.w, o.w, s2[4 * p + 3]);
                }
            }
    }
    if (SUMS) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float a = s1[i], b = s2[i];
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
            if (g == 0) {
                const int col = 16 * (i >> 2) + 4 * t + (i & 3);
                atomicAdd(red + col, a);
                atomicAdd(red + kMmaFout + col, b);
            }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x < 2 * kMmaFout) atomicAdd(bn_sums + threadIdx.x, (double)red[threadIdx.x]);
    }
}

template <bool SUMS, bool POOL>
__global__ void __maxnreg__(112)
rowgemm_mma_pipe_kernel(const float* __restrict__ x0, const float* __restrict__ xk, const float* __restrict__ W,
                        float* __restrict__ Y, double* __restrict__ bn_sums, long long R, int K, float* __restrict__ Ypool)
{
    rowgemm_mma_pipe_body<SUMS, 2, 4, POOL>(x0, xk, W, Y, bn_sums, R, K, Ypool);
}

template <bool SUMS, bool POOL>
__global__ void __launch_bounds__(kMpThreads, 1)
rowgemm_mma_pipe64_kernel(const float* __restrict__ x0, const float* __restrict__ xk, const float* __restrict__ W,
                          float* __restrict__ Y, double* __restrict__ bn_sums, long long R, int K, float* __restrict__ Ypool)
{
    rowgemm_mma_pipe_body<SUMS, 4, 5, POOL>(x0, xk, W, Y, bn_sums, R, K, Ypool);
}

// ---- weight gradient of the same narrow layer:  dW[fi*K + k, fo] += sum_r X_k[r, fi] dY[r, fo]   (autodiff of models.py:1077).
// The reduction runs over the ROWS, so in mma.m16n8k8 terms M = the 16 input features of a plane, N = the 32 output features
// (4 column tiles), k = 8 rows per instruction.  One CTA per SM (the K x 16 accumulator registers of a warp need the whole
// register file): 3 producer warps bring 128-row stages -- K plane tiles of X and the dY tile, all contiguous in memory -- by
// 1-D bulk copies into a 3-stage ring; each of the 8 consumer warps owns 16 rows of every stage for ALL planes (the dY
// fragments are read once per row step and shared by the K planes), splits both operands hi / lo in registers (3xTF32) and keeps
// its partial dW in registers until the end: warps -> shared memory -> one red.global.add per element and CTA.
// Fragment permutations: A's m index g / g+8 <-> features 4(g/2) + g%2 / + 2 (a lane reads one float4 of a row and uses two of
// its components), B's n index v of column tile j <-> output feature 4v + j (a lane reads one float4 of dY and uses component j
// for tile j), k index t / t+4 <-> rows t / t+4 of the 8-row step.
constexpr int kWgRows = 128, kWgStages = 3, kWgConsumers = 8, kWgProducers = 3;
constexpr int kWgThreads = (kWgConsumers + kWgProducers) * 32;

template <int K>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_mma_pipe_kernel(const float* __restrict__ x0, const float* __restrict__ xk, const float* __restrict__ dY,
                      float* __restrict__ dW, long long R)
{
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int kXTile = kWgRows * kMmaFin * 4, kYTile = kWgRows * kMmaFout * 4, kStage = K * kXTile + kYTile;
    float* sred = reinterpret_cast<float*>(smem + kWgStages * kStage);                  // [K*16][32]
    uint64_t* full = reinterpret_cast<uint64_t*>(sred + K * kMmaFin * kMmaFout);
    uint64_t* empty = full + kWgStages;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = lane & 3, g = lane >> 2;
    for (int i = threadIdx.x; i < K * kMmaFin * kMmaFout; i += kWgThreads) sred[i] = 0.f;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kWgStages; ++s) { mp_bar_init(full + s, kWgProducers); mp_bar_init(empty + s, kWgConsumers); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long plane = R * kMmaFin;
    const long long nblk = (R + kWgRows - 1) / kWgRows;
    const long long b0 = nblk * blockIdx.x / gridDim.x, b1 = nblk * (blockIdx.x + 1) / gridDim.x;
    const int nst = (int)(b1 - b0);

    if (warp >= kWgConsumers) {
        const int pw = warp - kWgConsumers;
        if (lane == 0) {
            for (int it = 0; it < nst; ++it) {
                const int s = it % kWgStages, u = it / kWgStages;
                if (u > 0) mp_bar_wait(empty + s, (uint32_t)(u - 1) & 1u);
                const long long row0 = (b0 + it) * kWgRows;
                const uint32_t rows = (uint32_t)(R - row0 < kWgRows ? R - row0 : kWgRows);
                unsigned char* st = smem + (size_t)s * kStage;
                uint32_t bytes = 0;
                for (int c = pw; c <= K; c += kWgProducers) bytes += rows * (c < K ? kMmaFin : kMmaFout) * 4;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mp_bar_expect_tx(full + s, bytes);
                for (int c = pw; c <= K; c += kWgProducers) {
                    if (c < K) mp_bulk_g2s(st + c * kXTile, (c == 0 ? x0 : xk + (long long)(c - 1) * plane) + row0 * kMmaFin, rows * kMmaFin * 4, full + s);
                    else mp_bulk_g2s(st + K * kXTile, dY + row0 * kMmaFout, rows * kMmaFout * 4, full + s);
                }
            }
        }
        return;
    }

    float acc[K][4][4];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[k][j][e] = 0.f;
    const bool odd = (g & 1) != 0;
    for (int it = 0; it < nst; ++it) {
        const int s = it % kWgStages, u = it / kWgStages;
        const long long row0 = (b0 + it) * kWgRows;
        const unsigned char* st = smem + (size_t)s * kStage;
        const float4* Ys = reinterpret_cast<const float4*>(st + K * kXTile);            // [128][8]
        mp_bar_wait(full + s, (uint32_t)u & 1u);
#pragma unroll
        for (int step = 0; step < 2; ++step) {
            const int r0 = warp * 16 + step * 8 + t, r1 = r0 + 4;                       // rows of this lane in the stage
            const bool ok0 = row0 + r0 < R, ok1 = row0 + r1 < R;                        // (stale shared memory past R)
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 y0 = ok0 ? Ys[r0 * 8 + g] : zero4, y1 = ok1 ? Ys[r1 * 8 + g] : zero4;
            const float yv0[4] = {y0.x, y0.y, y0.z, y0.w}, yv1[4] = {y1.x, y1.y, y1.z, y1.w};
            uint32_t bh[4][2], bl[4][2];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t h0 = __float_as_uint(yv0[j]) & 0xffffe000u, h1 = __float_as_uint(yv1[j]) & 0xffffe000u;
                bh[j][0] = h0; bh[j][1] = h1;
                bl[j][0] = __float_as_uint(yv0[j] - __uint_as_float(h0));
                bl[j][1] = __float_as_uint(yv1[j] - __uint_as_float(h1));
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float4* Xs = reinterpret_cast<const float4*>(st + k * kXTile);    // [128][4]
                const float4 a0 = ok0 ? Xs[r0 * 4 + (g >> 1)] : zero4, a1 = ok1 ? Xs[r1 * 4 + (g >> 1)] : zero4;
                const float f[4] = {odd ? a0.y : a0.x, odd ? a0.w : a0.z, odd ? a1.y : a1.x, odd ? a1.w : a1.z};
                uint32_t ah[4], al[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t hb = __float_as_uint(f[e]) & 0xffffe000u;
                    ah[e] = hb;
                    al[e] = __float_as_uint(f[e] - __uint_as_float(hb));
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    mma_tf32_16x8x8(acc[k][j], al, bh[j][0], bh[j][1]);
                    mma_tf32_16x8x8(acc[k][j], ah, bl[j][0], bl[j][1]);
                    mma_tf32_16x8x8(acc[k][j], ah, bh[j][0], bh[j][1]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mp_bar_arrive(empty + s);
    }
    // partial dW of this warp -> shared memory -> global
    const int fA = 4 * (g >> 1) + (g & 1);
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int fi = fA + ((e & 2) ? 2 : 0);                                   // c2, c3: m index g + 8
                const int fo = 4 * (2 * t + (e & 1)) + j;
                atomicAdd(sred + (fi * K + k) * kMmaFout + fo, acc[k][j][e]);
            }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    for (int i = threadIdx.x; i < K * kMmaFin * kMmaFout; i += kWgConsumers * 32)
        asm volatile("red.global.add.f32 [%0], %1;" :: "l"(dW + i), "f"(sred[i]) : "memory");
}

template <int K>
static int wgrad_mma_launch(const float* x0, const float* xk, const float* dY, float* dW, long long R, cudaStream_t s)
{
    const long long nblk = (R + kWgRows - 1) / kWgRows;
    const int grid = (int)(nblk < kNumSMs ? nblk : kNumSMs);
    const size_t smem = (size_t)kWgStages * (K * kWgRows * kMmaFin * 4 + kWgRows * kMmaFout * 4) + (size_t)K * kMmaFin * kMmaFout * 4
                        + 2 * kWgStages * 8 + 64;
    static bool cfg = false;
    if (!cfg) {
        cudaError_t ce = cudaFuncSetAttribute(wgrad_mma_pipe_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (ce != cudaSuccess) { set_error(cudaGetErrorString(ce)); return (int)ce; }
        cfg = true;
    }
    wgrad_mma_pipe_kernel<K><<<grid, kWgThreads, smem, s>>>(x0, xk, dY, dW, R);
    DSB_LAUNCH_CHECK();
}

// returns 1 when the shape does not qualify
int contract_wgrad_mma(const float* x0, const float* xk, const float* dY, float* dW, long long R, int Fin, int K, int Fout,
                       cudaStream_t s)
{
    // Wider layers (Fin 32 / 64 -> Fout 64) were tried with the consumer warps split over the output features (191 us at cosmo
    // layer 3 against 70 us): with row pitches of 128 / 256 bytes the four rows a fragment load touches fall on the same banks
    // (4-way conflicts on every operand read) -- they need a swizzled staging, which a 1-D bulk copy cannot write.
    if (Fin != kMmaFin || Fout != kMmaFout || R < kWgRows) return 1;
    if ((((uintptr_t)x0 | (uintptr_t)xk | (uintptr_t)dY) & 15) != 0) return 1;
    if (K == 5) return wgrad_mma_launch<5>(x0, xk, dY, dW, R, s);
    if (K == 4) return wgrad_mma_launch<4>(x0, xk, dY, dW, R, s);
    if (K == 3) return wgrad_mma_launch<3>(x0, xk, dY, dW, R, s);
    return 1;
}

// returns 1 when the shape does not qualify
int contract_fwd_mma(const float* x0, const float* xk, const float* W, float* Y, long long R, int Fin, int K, int Fout,
                     double* bn_sums, cudaStream_t s, float* Ypool)
{
    if (Fin != kMmaFin || Fout != kMmaFout || K < 1 || K > kMmaMaxK) return 1;
    if ((((uintptr_t)x0 | (uintptr_t)xk | (uintptr_t)Y | (uintptr_t)Ypool) & 15) != 0) return 1;
    if (Ypool && (R & 3)) return 1;
    const long long cap = 2LL * kNumSMs;
    int variant = DSB_ENV_INT("DSB200_FWD_MMA", 3);                  // 3 (default): 64 rows per warp, one CTA per SM (100.6 us with
                                                                     // the sums at cosmo layer 2); 1: 32 rows, two CTAs (105.8 us)
    if (Ypool && variant == 2) variant = 3;                          // the register experiment has no pooled epilogue
    if (variant == 2) {                                              // register version (experiment)
        const long long nblk = (R + 31) / 32;
        const int grid = (int)((nblk + 7) / 8 < cap ? (nblk + 7) / 8 : cap);
        rowgemm_mma_kernel<<<grid, 256, 0, s>>>(x0, xk, W, Y, bn_sums, R, K);
        DSB_LAUNCH_CHECK();
    }
    const size_t fixed = (size_t)kMmaMaxK * 2 * 4 * 32 * 16 + 2 * kMmaFout * 4 + 2 * 8 * 8 + 64;
    if (variant == 3) {                                              // 64 rows per warp, one CTA per SM
        const long long nb64 = (R + 511) / 512;
        const int grid64 = (int)(nb64 < kNumSMs ? nb64 : kNumSMs);
        const size_t smem64 = (size_t)5 * 512 * kMmaFin * 4 + fixed;
        static bool cfg64 = false;
        if (!cfg64) {
            cudaError_t ce = cudaFuncSetAttribute(rowgemm_mma_pipe64_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (ce == cudaSuccess) ce = cudaFuncSetAttribute(rowgemm_mma_pipe64_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (ce == cudaSuccess) ce = cudaFuncSetAttribute(rowgemm_mma_pipe64_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (ce == cudaSuccess) ce = cudaFuncSetAttribute(rowgemm_mma_pipe64_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (ce != cudaSuccess) { set_error(cudaGetErrorString(ce)); return (int)ce; }
            cfg64 = true;
        }
        if (Ypool) {
            if (bn_sums) rowgemm_mma_pipe64_kernel<true, true><<<grid64, kMpThreads, smem64, s>>>(x0, xk, W, Y, bn_sums, R, K, Ypool);
            else rowgemm_mma_pipe64_kernel<false, true><<<grid64, kMpThreads, smem64, s>>>(x0, xk, W, Y, bn_sums, R, K, Ypool);
        } else if (bn_sums) rowgemm_mma_pipe64_kernel<true, false><<<grid64, kMpThreads, smem64, s>>>(x0, xk, W, Y, bn_sums, R, K, Ypool);
        else rowgemm_mma_pipe64_kernel<false, false><<<grid64, kMpThreads, smem64, s>>>(x0, xk, W, Y, bn_sums, R, K, Ypool);
        DSB_LAUNCH_CHECK();
    }
    const long long nblk = (R + 255) / 256;
    const int grid = (int)(nblk < cap ? nblk : cap);
    const size_t smem = (size_t)4 * 256 * kMmaFin * 4 + fixed;
    static bool cfg = false;
    if (!cfg) {
        cudaError_t ce = cudaFuncSetAttribute(rowgemm_mma_pipe_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(rowgemm_mma_pipe_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(rowgemm_mma_pipe_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(rowgemm_mma_pipe_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
        if (ce != cudaSuccess) { set_error(cudaGetErrorString(ce)); return (int)ce; }
        cfg = true;
    }
    if (Ypool) {
        if (bn_sums) rowgemm_mma_pipe_kernel<true, true><<<grid, kMpThreads, smem, s>>>(x0, xk, W, Y, bn_sums, R, K, Ypool);
        else rowgemm_mma_pipe_kernel<false, true><<<grid, kMpThreads, smem, s>>>(x0, xk, W, Y, bn_sums, R, K, Ypool);
    } else if (bn_sums) rowgemm_mma_pipe_kernel<true, false><<<grid, kMpThreads, smem, s>>>(x0, xk, W, Y, bn_sums, R, K, Ypool);
    else rowgemm_mma_pipe_kernel<false, false><<<grid, kMpThreads, smem, s>>>(x0, xk, W, Y, bn_sums, R, K, Ypool);
    DSB_LAUNCH_CHECK();
}

}  // namespace dsb200
