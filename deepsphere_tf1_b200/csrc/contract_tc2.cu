// K2 forward with the ROWS on the N side of the tensor-core instruction (reference models.py:1062-1078):
//   Y^T[fo, r] = sum_{k, fi} W[fi*K + k, fo] * X_k[r, fi]
// One tcgen05.mma kind::tf32 occupies the tensor pipe ~87 ns whatever its M and N (tools/experiments/mma_rate.cu),
// so the row-GEMM of contract_tc.cu (M = 128 rows, N = 2*Fout) is tensor-issue bound: 2 instructions per 8 reduction
// elements and 128 rows.  Here one instruction covers 256 rows:
//   A operand (M = 128, resident in shared memory): the weights with hi / lo interleaved on M, row 2*fo = W_hi[:, fo],
//                                                   row 2*fo + 1 = W_lo[:, fo], zero rows above 2*Fout
//   B operand (N = 256): the landed activation tile [256 rows x kc], K-major -- the same TMA / swizzle layout as before
//   D[2*fo + part, r] += A . X      (X read truncated to TF32 by the tensor core)
//   D                 += A . X_lo   (X_lo = the bits the truncation drops; the extra W_lo X_lo term is below fp32 resolution)
// i.e. 2 instructions per 8 reduction elements and 256 rows: half the tensor time.  Epilogue: lanes 2*fo and 2*fo + 1 of
// a warp are added with one shuffle, the batch-norm sums are plain per-lane accumulations over the columns, and a store
// instruction writes two rows of 16 consecutive features.
#include "common.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

namespace dsb200 {
namespace tc {

constexpr int kTileN = 256;                 // rows per tile (the N of the instruction)
constexpr int kImgRows = 128;               // the M of the instruction
constexpr int kCgThreads = 320;             // warp 0 TMA, warp 1 MMA, warps 2-5 lo split, warps 6-9 epilogue
constexpr int kCgStagesMax = 8;

struct alignas(64) ColGemmParams {
    CUtensorMap tmA0, tmAk;    // plane 0 / planes 1..K-1, boxes of 256 rows x kc columns
    const float* wimg;         // weight image [nkb][128][kc], K-major swizzled (colgemm_pack_kernel)
    float* out;                // Y [R, Fout] (may be NULL when only the pooled output is wanted)
    float* pool_out;           // may be NULL: max over every 4 consecutive rows of Y, [R/4, Fout]
    double* bn_sums;           // may be NULL
    long long R;
    int Fin, K, Fout, kc, nchunk, nkb, swz, nstages, ntiles;
    int acc_stages;            // 2: double-buffered accumulators (512 TMEM columns, one CTA per SM); 1: two CTAs per SM
};

__global__ void __launch_bounds__(256)
colgemm_pack_kernel(const float* __restrict__ W, float* __restrict__ image, int Fin, int K, int Fout, int kc, int nchunk, int nkb, int swz)
{
    const int per_blk = kImgRows * kc;
    const int total = nkb * per_blk;
    const int rowbytes = kc * 4;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int j = idx / per_blk;
        const int rem = idx - j * per_blk;
        const int n2 = rem / kc, kk = rem - n2 * kc;
        const int fo = n2 >> 1, part = n2 & 1;
        float v = 0.f;
        if (fo < Fout) {
            const int k = j / nchunk, c = j - k * nchunk;
            const float w = __ldg(W + ((long long)(c * kc + kk) * K + k) * Fout + fo);
            const float hi = to_tf32(w);
            v = part ? to_tf32(w - hi) : hi;
        }
        const int chunk = (kk >> 2) ^ swizzle_xor(n2, swz);
        image[((size_t)j * kImgRows * rowbytes + (size_t)n2 * rowbytes + chunk * 16 + (kk & 3) * 4) >> 2] = v;
    }
}

template <bool SUMS, bool POOL>
__global__ void __launch_bounds__(kCgThreads, 2)
colgemm_tc_kernel(const __grid_constant__ ColGemmParams p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rowbytes = p.kc * 4;
    const uint32_t tile_bytes = (uint32_t)kTileN * rowbytes;          // one k-block of activations
    const uint32_t wblk_bytes = (uint32_t)kImgRows * rowbytes;
    const uint32_t stage_bytes = 2u * tile_bytes;                     // [X][X_lo]
    const int S = p.nstages;

    uint8_t* sW = smem;
    uint8_t* sA = sW + (size_t)p.nkb * wblk_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)S * stage_bytes);
    uint64_t* full_tma = bars;
    uint64_t* full_mma = bars + kCgStagesMax;
    uint64_t* empty = bars + 2 * kCgStagesMax;
    uint64_t* acc_full = bars + 3 * kCgStagesMax;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full_tma + s, 1); mbar_init(full_mma + s, 128); mbar_init(empty + s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(acc_full + a, 1); mbar_init(acc_empty + a, 128); }
        fence_barrier_init();
        prefetch_tmap(&p.tmA0);
        if (p.K > 1) prefetch_tmap(&p.tmAk);
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)(p.acc_stages * kTileN));
    {
        const int nvec = (int)((size_t)p.nkb * wblk_bytes / 16);
        const float4* src = reinterpret_cast<const float4*>(p.wimg);
        float4* dst = reinterpret_cast<float4*>(sW);
        for (int i = threadIdx.x; i < nvec; i += kCgThreads) dst[i] = __ldg(src + i);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ======================= TMA producer: one k-block (256 rows x kc) per stage =======================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
                const int row0 = t * kTileN;
                int k = 0, c = 0;
                for (int j = 0; j < p.nkb; ++j) {
                    mbar_wait(empty + s, ph ^ 1u, 21);
                    mbar_expect_tx(full_tma + s, tile_bytes);
                    uint8_t* dst = sA + (size_t)s * stage_bytes;
                    if (k == 0) tma_load_2d(dst, &p.tmA0, full_tma + s, c * p.kc, row0);
                    else tma_load_3d(dst, &p.tmAk, full_tma + s, c * p.kc, row0, k - 1);
                    if (++c == p.nchunk) { c = 0; ++k; }
                    if (++s == S) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(kImgRows, kTileN, 0, 0);
            const uint32_t sbo = 8u * rowbytes;
            const int ksteps = p.kc / 8;
            const uint64_t dX0 = make_smem_desc(smem_u32(sA), 16, sbo, p.swz);
            const uint64_t dW0 = make_smem_desc(smem_u32(sW), 16, sbo, p.swz);
            const uint32_t x_stage = stage_bytes >> 4, x_lo = tile_bytes >> 4, w_blk = wblk_bytes >> 4;
            int s = 0, ti = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
                const int a = p.acc_stages == 2 ? (ti & 1) : 0;
                const uint32_t use = (uint32_t)(p.acc_stages == 2 ? (ti >> 1) : ti);
                mbar_wait(acc_empty + a, (use & 1u) ^ 1u, 22);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(a * kTileN);
                for (int j = 0; j < p.nkb; ++j) {
                    mbar_wait(full_mma + s, ph, 23);
                    tc_fence_after();
                    const uint64_t dX = dX0 + (uint64_t)(x_stage * (uint32_t)s);
                    const uint64_t dW = dW0 + (uint64_t)(w_blk * (uint32_t)j);
#pragma unroll 4
                    for (int ks = 0; ks < ksteps; ++ks) {
                        mma_tf32(d, dW + 2u * ks, dX + 2u * ks, idesc, (j > 0 || ks > 0) ? 1u : 0u);     // [W_hi; W_lo] . X
                        mma_tf32(d, dW + 2u * ks, dX + x_lo + 2u * ks, idesc, 1u);                       // [W_hi; W_lo] . X_lo
                    }
                    tc_commit(empty + s);
                    if (++s == S) { s = 0; ph ^= 1u; }
                }
                tc_commit(acc_full + a);
            }
        }
    } else if (warp < 6) {
        // ======================= lo part of every landed tile (the tile itself is the hi part) =======================
        const int tt = threadIdx.x - 64;
        int s = 0;
        uint32_t ph = 0;
        const int nvec = (int)(tile_bytes / 16);
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            for (int j = 0; j < p.nkb; ++j) {
                mbar_wait(full_tma + s, ph, 24);
                const float4* X = reinterpret_cast<const float4*>(sA + (size_t)s * stage_bytes);
                float4* Xl = reinterpret_cast<float4*>(sA + (size_t)s * stage_bytes + tile_bytes);
#pragma unroll 4
                for (int i = tt; i < nvec; i += 128) Xl[i] = tf32_dropped_bits(X[i]);
                fence_proxy_async();
                mbar_arrive(full_mma + s);
                if (++s == S) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ======================= epilogue: lanes (2*fo, 2*fo+1) = (hi, lo) rows of feature fo =======================
        const int q = warp & 3;                                      // TMEM lane quarter == image rows 32q .. 32q+31
        const int fo = q * 16 + (lane >> 1);
        const int odd = lane & 1;
        const bool live = q * 16 < p.Fout;                           // this warp's image rows hold real features
        float s1 = 0.f, s2 = 0.f;
        int ti = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
            const int a = p.acc_stages == 2 ? (ti & 1) : 0;
            const uint32_t use = (uint32_t)(p.acc_stages == 2 ? (ti >> 1) : ti);
            mbar_wait(acc_full + a, use & 1u, 25);
            tc_fence_after();
            if (live) {
                const uint32_t tbase = tmem_base + (uint32_t)(a * kTileN) + ((uint32_t)(q * 32) << 16);
                const long long row0 = (long long)t * kTileN;
                for (int c0 = 0; c0 < kTileN; c0 += 32) {
                    float v[32];
                    tmem_ld16(tbase + c0, v);
                    tmem_ld16(tbase + c0 + 16, v + 16);
                    tmem_ld_wait();
                    float pm = 0.f;
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        // hi + lo of both rows; the even lane keeps row c0+i, the odd lane row c0+i+1
                        const float e = v[i] + __shfl_xor_sync(0xffffffffu, v[i], 1);
                        const float o = v[i + 1] + __shfl_xor_sync(0xffffffffu, v[i + 1], 1);
                        const float mine = odd ? o : e;
                        const long long r = row0 + c0 + i + odd;
                        if (fo < p.Fout && r < p.R) {
                            if (!POOL || p.out) p.out[r * p.Fout + fo] = mine;
                            if (SUMS) { s1 += mine; s2 = fmaf(mine, mine, s2); }
                        }
                        if (POOL) {
                            // max-pool over the 4 consecutive rows c0+i-2 .. c0+i+1 (i = 2 mod 4): this lane holds two of them (in
                            // `pm` and `mine`), its partner the other two; the pooling commutes with the batch norm / bias /
                            // activation that follow (all non-decreasing per feature, models.py:1235-1243)
                            if ((i & 2) == 0) pm = mine;
                            else {
                                float m4 = fmaxf(pm, mine);
                                m4 = fmaxf(m4, __shfl_xor_sync(0xffffffffu, m4, 1));
                                const long long rg = row0 + c0 + i - 2;
                                if (!odd && fo < p.Fout && rg < p.R) p.pool_out[(rg >> 2) * p.Fout + fo] = m4;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(acc_empty + a);
        }
        if (SUMS && live) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
            s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
            if (!odd && fo < p.Fout) { atomicAdd(p.bn_sums + fo, (double)s1); atomicAdd(p.bn_sums + p.Fout + fo, (double)s2); }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)(p.acc_stages * kTileN));
}

size_t colgemm_workspace_bytes(int Fin, int K) { return (size_t)K * Fin * kImgRows * 4 + 1024; }

// Returns 1 when the shape is not supported (the caller uses the row-GEMM).
int contract_fwd_tc_cols(const float* x0, const float* xk, const float* W, float* Y, long long R, int Fin, int K, int Fout,
                         double* bn_sums, void* workspace, cudaStream_t s, float* Ypool)
{
    if (DSB_ENV_INT("DSB200_NO_COLGEMM", 0)) return 1;
    if (Ypool && (R & 3)) return 1;
    if (!workspace || R < kTileN || R > 0x7fffff00LL || Fout > 64 || (Fout & 3) || (Fin & 15)) return 1;
    // Measured (tools/sweep_gemm.py): 77 vs 90 us at cosmo layer 3 (Fin = 32), but 181 vs 136 us at layer 2 (Fin = 16, Fout = 32:
    // half of the 128 accumulator lanes and half of the epilogue warps idle, and TMA delivers 64-byte rows no faster than ~4 ns
    // each per SM).  DSB200_CG_ALL=1 forces this kernel wherever the shape fits.
    if (Fin < 32 && !DSB_ENV_INT("DSB200_CG_ALL", 0)) return 1;
    if ((((uintptr_t)x0 | (uintptr_t)xk | (uintptr_t)Y | (uintptr_t)workspace) & 15) != 0) return 1;
    ColGemmParams p;
    p.kc = 16; p.swz = 2;
    p.nchunk = Fin / p.kc;
    p.nkb = K * p.nchunk;
    const size_t wbytes = (size_t)p.nkb * kImgRows * p.kc * 4;
    const size_t stage = 2 * (size_t)kTileN * p.kc * 4;
    const size_t fixed = 1024 + 3 * kCgStagesMax * 8 + 4 * 8 + 16 + 64;
    int ctas = 1;
    { const int e = DSB_ENV_INT("DSB200_CG_CTAS", 0); if (e) ctas = e == 2 ? 2 : 1; }
    size_t budget = (size_t)(227 * 1024) / ctas - (ctas > 1 ? 1024 : 0);
    if (ctas == 2 && wbytes + fixed + 2 * stage > budget) { ctas = 1; budget = 227 * 1024; }
    if (wbytes + fixed + (ctas == 2 ? 2 : 3) * stage > budget) return 1;
    size_t ns = (budget - wbytes - fixed) / stage;
    if (ns > kCgStagesMax) ns = kCgStagesMax;
    p.nstages = (int)ns;
    p.acc_stages = ctas == 2 ? 1 : 2;
    size_t smem = wbytes + fixed + ns * stage;
    const size_t min_smem = (size_t)(227 * 1024) / (ctas + 1) + 1024;
    if (smem < min_smem) smem = min_smem;
    if (make_tmap_f32(&p.tmA0, x0, Fin, R, 1, p.kc, kTileN, p.swz, false)) return 1;
    if (K > 1) { if (make_tmap_f32(&p.tmAk, xk, Fin, R, K - 1, p.kc, kTileN, p.swz, true)) return 1; }
    else p.tmAk = p.tmA0;
    p.wimg = (const float*)workspace; p.out = Y; p.pool_out = Ypool; p.bn_sums = bn_sums; p.R = R; p.Fin = Fin; p.K = K; p.Fout = Fout;
    p.ntiles = (int)((R + kTileN - 1) / kTileN);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(colgemm_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(colgemm_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(colgemm_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(colgemm_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    const int total = p.nkb * kImgRows * p.kc;
    colgemm_pack_kernel<<<(total + 255) / 256, 256, 0, s>>>(W, (float*)workspace, Fin, K, Fout, p.kc, p.nchunk, p.nkb, p.swz);
    count_launch();
    const int slots = kNumSMs * ctas;
    const int grid = p.ntiles < slots ? p.ntiles : slots;
    if (Ypool) {
        if (bn_sums) colgemm_tc_kernel<true, true><<<grid, kCgThreads, smem, s>>>(p);
        else colgemm_tc_kernel<false, true><<<grid, kCgThreads, smem, s>>>(p);
    } else if (bn_sums) colgemm_tc_kernel<true, false><<<grid, kCgThreads, smem, s>>>(p);
    else colgemm_tc_kernel<false, false><<<grid, kCgThreads, smem, s>>>(p);
    DSB_LAUNCH_CHECK();
}

}  // namespace tc
}  // namespace dsb200
