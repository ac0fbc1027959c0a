// K2 on the 5th-generation tensor cores: the dense contraction of cgcnn.chebyshev5
//   Y[r, fo] = sum_{k, fi} X_k[r, fi] * W[fi*K + k, fo]            (reference models.py:1062-1078)
// and its data gradient  dX_k[r, fi] = sum_fo dY[r, fo] * W[fi*K + k, fo]  (autodiff of 1077),
// as persistent, warp-specialised tcgen05 kernels:
//   warp 0      TMA producer: 128-row x kc-column fp32 tiles of the activation planes -> shared memory
//               (cp.async.bulk.tensor, hardware 32/64/128-byte swizzle, mbarrier completion)
//   warps 2-5   split every landed tile into TF32 "hi" (in place) and "lo" (second buffer) parts
//   warp 1      one thread issues tcgen05.mma kind::tf32 into TMEM:  D = A_hi*[W_hi | W_lo] + A_lo*W_hi
//               (3xTF32 error-compensated product: fp32-level accuracy, fp32 accumulation)
//   warps 6-9   epilogue: tcgen05.ld the accumulator, add the two halves, batch-norm sums
//               (tf.nn.moments, models.py:1197-1203), coalesced fp32 stores
// The weights are tiny: they are split and laid out once per CTA in the canonical K-major swizzled
// layout and stay resident in shared memory.  TMEM accumulators are double buffered when they fit.
#include "common.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

namespace dsb200 {
namespace tc {

EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int make_tmap_f32(CUtensorMap* out, const float* base, long long cols, long long rows, long long planes,
                  int box_cols, int box_rows, int swz, bool rank3)
{
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return -1;
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)cols * 4, (cuuint64_t)cols * 4 * (cuuint64_t)rows};
    cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapSwizzle sw = swz == 4 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                : swz == 3 ? CU_TENSOR_MAP_SWIZZLE_128B : swz == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swz == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    const cuuint32_t rank = rank3 ? 3 : 2;
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, (void*)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

// general fp32 tensor map: dims / box innermost first, strides (bytes) for dims 1..rank-1
static int make_tmap_nd(CUtensorMap* out, const float* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                        const cuuint32_t* box, CUtensorMapSwizzle sw)
{
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return -1;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void*)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

constexpr int kTileM = 128;
constexpr int kMaxStages = 8;
enum { MODE_FWD = 0, MODE_BWD_DATA = 1 };

// warp roles of the row-GEMM kernel
constexpr int kRgThreads = 448;          // 14 warps
constexpr int kWarpTma = 0, kWarpMma = 1, kWarpSplit0 = 2, kWarpEpi0 = 10;   // split: warps 2-5 and 6-9

struct alignas(64) RowGemmParams {
    CUtensorMap tmA0;          // plane 0 (x0, or dY for the data gradient)
    CUtensorMap tmAk;          // planes 1..K-1 (forward only)
    CUtensorMap tmOut;         // (unused: a tensor-store epilogue was row-rate bound)
    int epi_tma, epi_cols;     // data gradient: stage 128 rows x epi_cols (= Fin) of a plane in shared memory, 1-D bulk store
    const float* W;            // [Fin*K, Fout], row = fi*K + k
    const float* wpack;        // pre-split weights in the shared-memory image layout (pack_w_kernel), or NULL
    float* out;                // Y [R, Fout] (forward: may be NULL when only the pooled output is wanted)  |  G [K][R][Fin]
    float* pool_out;           // forward only, may be NULL: max over every 4 consecutive rows of Y, [R/4, Fout]
    double* bn_sums;           // forward only, may be NULL
    long long R;
    int Fin, K, Fout;
    int kc, nchunk, nkb;       // reduction: chunk width, chunks per plane, k-blocks per tile
    int G;                     // k-blocks per pipeline stage
    int Npad, Nreal, n0;       // accumulator width (multiple of 16), valid columns, first column of this n-block
    int swz;                   // 1 / 2 / 3 = 32 / 64 / 128-byte swizzle (row of a tile = kc*4 bytes)
    int nstages, acc_stages;
    int ntiles;
    int wstream;               // 1: the weight image is too large to stay resident: each stage carries its k-blocks' weights
    int tmem_cols;             // TMEM columns to allocate (power of two >= acc_stages * 2 * Npad)
    int nostore;               // experiment: skip the epilogue stores (DSB200_RG_NOSTORE)
    long long* dbg;            // experiment: per-stage timestamps of CTA 0 (DSB200_RG_DBG = device pointer)
};

// Element (j, n2, kk) of the weight image: k-block j, row n2 in [0, 2*Npad) (hi rows then lo rows), column kk.
template <int MODE>
__device__ __forceinline__ float w_image_value(const RowGemmParams& p, int j, int n2, int kk)
{
    const int half = n2 >= p.Npad;
    const int n = n2 - half * p.Npad;
    float w = 0.f;
    if (n < p.Nreal) {
        if (MODE == MODE_FWD) {
            const int k = j / p.nchunk, c = j - k * p.nchunk;
            w = __ldg(p.W + ((long long)(c * p.kc + kk) * p.K + k) * p.Fout + (p.n0 + n));
        } else {
            const int q = p.n0 + n;
            const int kq = q / p.Fin, fi = q - kq * p.Fin;
            w = __ldg(p.W + ((long long)fi * p.K + kq) * p.Fout + (j * p.kc + kk));
        }
    }
    const float hi = to_tf32(w);
    return half ? to_tf32(w - hi) : hi;
}
__device__ __forceinline__ uint32_t w_image_offset(const RowGemmParams& p, int j, int n2, int kk)
{
    const int rowbytes = p.kc * 4;
    const int chunk = (kk >> 2) ^ swizzle_xor(n2, p.swz);
    return (uint32_t)j * (2u * p.Npad * rowbytes) + (uint32_t)n2 * rowbytes + chunk * 16 + (kk & 3) * 4;
}

// Weights -> TF32 hi / lo, canonical K-major swizzled layout, once per launch into global scratch.
template <int MODE>
__global__ void __launch_bounds__(256) pack_w_kernel(const __grid_constant__ RowGemmParams p, float* __restrict__ image)
{
    const int per_blk = 2 * p.Npad * p.kc;
    const int total = p.nkb * per_blk;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int j = idx / per_blk;
        const int rem = idx - j * per_blk;
        const int n2 = rem / p.kc, kk = rem - n2 * p.kc;
        image[w_image_offset(p, j, n2, kk) >> 2] = w_image_value<MODE>(p, j, n2, kk);
    }
}

// Column sums over the 32 rows held by the lanes of a warp: v[i] = value of column i in this lane's row.
// Transposing butterfly (31 shuffles): afterwards lane l holds the sum of column l in v[0].
__device__ __forceinline__ void warp_column_sums(float (&v)[32], int lane)
{
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float keep = up ? v[i + s] : v[i];
            const float send = up ? v[i] : v[i + s];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
}

__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define RG_DBG(slot, ev) do { if (p.dbg && blockIdx.x == 0 && (slot) < 256) p.dbg[(slot) * 8 + (ev)] = gtimer(); } while (0)

// POOL: forward only, the epilogue also emits the max over every 4 consecutive rows (a separate instance, so that the training
// kernels -- which keep the pooling in the batch-norm pass, see models.py -- are compiled exactly as before)
template <int MODE, bool SUMS, bool POOL = false>
__global__ void __launch_bounds__(kRgThreads, 2)
rowgemm_tc_kernel(const __grid_constant__ RowGemmParams p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rowbytes = p.kc * 4;
    const uint32_t tile_bytes = (uint32_t)kTileM * rowbytes;
    const uint32_t wblk_bytes = 2u * p.Npad * rowbytes;
    // stage: [G tiles of A][G tiles of A_lo] (+ [G weight blocks] when the weights are streamed)
    const uint32_t stage_bytes = 2u * p.G * tile_bytes + (p.wstream ? p.G * wblk_bytes : 0u);
    const int S = p.nstages;
    const int ngroups = (p.nkb + p.G - 1) / p.G;                  // pipeline stages per row tile

    uint8_t* sW = smem;
    uint8_t* sA = sW + (p.wstream ? 0 : (size_t)p.nkb * wblk_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)S * stage_bytes);
    uint64_t* full_tma = bars;
    uint64_t* full_mma = bars + kMaxStages;
    uint64_t* empty = bars + 2 * kMaxStages;
    uint64_t* acc_full = bars + 3 * kMaxStages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* red = reinterpret_cast<float*>(tmem_slot + 2);           // [2][256] batch-norm partial sums
    // data-gradient epilogue staging: [K planes][128 rows][epi_cols] fp32
    uint8_t* epi = smem + ((((size_t)(reinterpret_cast<uint8_t*>(red + 2 * 256) - smem)) + 1023) & ~(size_t)1023);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full_tma + s, 1); mbar_init(full_mma + s, 128); mbar_init(empty + s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(acc_full + a, 1); mbar_init(acc_empty + a, 128); }
        fence_barrier_init();
        prefetch_tmap(&p.tmA0);
        if (MODE == MODE_FWD && p.K > 1) prefetch_tmap(&p.tmAk);
    }
    if (warp == kWarpMma) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    if (SUMS) for (int i = threadIdx.x; i < 2 * 256; i += kRgThreads) red[i] = 0.f;

    // ---- weights: rows [0,Npad) = TF32 hi, [Npad,2Npad) = lo, K-major swizzled, resident for the whole kernel
    if (p.wstream) {
        // nothing resident
    } else if (p.wpack) {
        const int nvec = (int)((size_t)p.nkb * wblk_bytes / 16);
        const float4* src = reinterpret_cast<const float4*>(p.wpack);
        float4* dst = reinterpret_cast<float4*>(sW);
        for (int i = threadIdx.x; i < nvec; i += kRgThreads) dst[i] = __ldg(src + i);
    } else {
        const int per_blk = 2 * p.Npad * p.kc;
        const int total = p.nkb * per_blk;
        for (int idx = threadIdx.x; idx < total; idx += kRgThreads) {
            const int j = idx / per_blk;
            const int rem = idx - j * per_blk;
            const int n2 = rem / p.kc, kk = rem - n2 * p.kc;
            *reinterpret_cast<float*>(sW + w_image_offset(p, j, n2, kk)) = w_image_value<MODE>(p, j, n2, kk);
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int acc_cols = 2 * p.Npad;

    if (warp == kWarpTma) {
        // ======================= TMA producer =======================
        if (lane == 0) {
            int s = 0, dbgi = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
                const int row0 = t * kTileM;
                int k = 0, c = 0;                                   // plane, chunk within the plane
                for (int j0 = 0; j0 < p.nkb; j0 += p.G) {
                    const int cnt = min(p.G, p.nkb - j0);
                    RG_DBG(dbgi, 0);
                    mbar_wait(empty + s, ph ^ 1u, 1);
                    RG_DBG(dbgi, 1);
                    ++dbgi;
                    mbar_expect_tx(full_tma + s, cnt * (tile_bytes + (p.wstream ? wblk_bytes : 0u)));
                    uint8_t* dst = sA + (size_t)s * stage_bytes;
                    if (p.wstream)
                        bulk_load(dst + 2 * (size_t)p.G * tile_bytes, reinterpret_cast<const uint8_t*>(p.wpack) + (size_t)j0 * wblk_bytes,
                                  cnt * wblk_bytes, full_tma + s);
                    for (int u = 0; u < cnt; ++u, dst += tile_bytes) {
                        if (MODE == MODE_FWD) {
                            if (k == 0) tma_load_2d(dst, &p.tmA0, full_tma + s, c * p.kc, row0);
                            else tma_load_3d(dst, &p.tmAk, full_tma + s, c * p.kc, row0, k - 1);
                            if (++c == p.nchunk) { c = 0; ++k; }
                        } else {
                            tma_load_2d(dst, &p.tmA0, full_tma + s, (j0 + u) * p.kc, row0);
                        }
                    }
                    if (++s == S) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == kWarpMma) {
        // ======================= MMA issuer =======================
        if (lane == 0) {
            const bool merged = acc_cols <= 256;
            const uint32_t idesc_wide = make_idesc_tf32(kTileM, merged ? acc_cols : p.Npad, 0, 0);
            const uint32_t idesc_half = make_idesc_tf32(kTileM, p.Npad, 0, 0);
            const uint32_t sbo = 8u * rowbytes;
            const int ksteps = p.kc / 8;
            // descriptors differ only in the 14-bit start-address field (bytes >> 4): build them once
            const uint64_t dA0 = make_smem_desc(smem_u32(sA), 16, sbo, p.swz);
            const uint64_t dB0 = make_smem_desc(smem_u32(sW), 16, sbo, p.swz);
            const uint32_t a_stage = stage_bytes >> 4, a_tile = tile_bytes >> 4, a_lo = (p.G * tile_bytes) >> 4;
            const uint32_t b_blk = wblk_bytes >> 4, b_lo = ((uint32_t)p.Npad * rowbytes) >> 4;
            int s = 0, ti = 0, dbgm = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
                const int a = p.acc_stages == 2 ? (ti & 1) : 0;
                const uint32_t use = (uint32_t)(p.acc_stages == 2 ? (ti >> 1) : ti);
                mbar_wait(acc_empty + a, (use & 1) ^ 1u, 2);
                tc_fence_after();
                const uint32_t d_hi = tmem_base + (uint32_t)(a * acc_cols);
                const uint32_t d_lo = d_hi + (uint32_t)p.Npad;
                uint64_t dB = dB0;
                for (int j0 = 0; j0 < p.nkb; j0 += p.G) {
                    const int cnt = min(p.G, p.nkb - j0);
                    mbar_wait(full_mma + s, ph, 3);
                    RG_DBG(dbgm, 4);
                    tc_fence_after();
                    uint64_t dA = dA0 + (uint64_t)(a_stage * (uint32_t)s);
                    if (p.wstream) dB = dA + 2u * p.G * a_tile;                      // this stage's weight blocks
                    for (int u = 0; u < cnt; ++u, dA += a_tile, dB += b_blk) {
#pragma unroll 4
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint32_t acc = (j0 + u > 0 || ks > 0) ? 1u : 0u;
                            const uint64_t dAH = dA + 2u * ks, dBH = dB + 2u * ks;  // +32 bytes per k-step of 8
                            if (merged) {
                                mma_tf32(d_hi, dAH, dBH, idesc_wide, acc);          // A_hi * [W_hi | W_lo]
                            } else {
                                mma_tf32(d_hi, dAH, dBH, idesc_half, acc);
                                mma_tf32(d_lo, dAH, dBH + b_lo, idesc_half, acc);
                            }
                            mma_tf32(d_hi, dAH + a_lo, dBH, idesc_half, 1u);        // A_lo * W_hi
                        }
                    }
                    tc_commit(empty + s);                                       // stage free once the MMAs have read it
                    RG_DBG(dbgm, 5);
                    ++dbgm;
                    if (++s == S) { s = 0; ph ^= 1u; }
                }
                tc_commit(acc_full + a);
            }
        }
    } else if (warp < kWarpEpi0) {
        // ======================= TF32 hi / lo split: two groups of 4 warps take alternate stages =======================
        const int grp = (warp - kWarpSplit0) >> 2;
        const int tt = (threadIdx.x - kWarpSplit0 * 32) & 127;
        int s = 0;
        uint32_t ph = 0, it = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            for (int j0 = 0; j0 < p.nkb; j0 += p.G, ++it) {
                // A stage always belongs to the same group (s & 1), and a group walks its stages in issue order.  Alternating on
                // the iteration counter instead lets, with an odd number of stages, the OTHER group wait for the next phase of a
                // stage whose current phase has not completed yet -- a parity wait that passes immediately (the wgrad deadlock).
                if ((s & 1) == grp) {
                    const int cnt = min(p.G, p.nkb - j0);
                    mbar_wait(full_tma + s, ph, 4);
                    if (tt == 0) RG_DBG((int)it, 2);
                    float4* A = reinterpret_cast<float4*>(sA + (size_t)s * stage_bytes);
                    float4* Al = reinterpret_cast<float4*>(sA + (size_t)s * stage_bytes + (size_t)p.G * tile_bytes);
                    const int nvec = (int)(cnt * tile_bytes / 16);
#pragma unroll 4
                    for (int i = tt; i < nvec; i += 128) {
                        // The tensor core reads a TF32 operand truncated (low 13 mantissa bits ignored), so the landed
                        // fp32 tile IS the hi part; lo = the bits it drops, exact in fp32 (cvt.rna.tf32 would cost ~10 SASS
                        // instructions per element and a second store).
                        Al[i] = tf32_dropped_bits(A[i]);
                    }
                    fence_proxy_async();
                    mbar_arrive(full_mma + s);
                    if (tt == 0) RG_DBG((int)it, 3);
                }
                if (++s == S) { s = 0; ph ^= 1u; }
            }
        }
        (void)ngroups;
    } else {
        // ======================= epilogue =======================
        const int q = warp & 3;                                     // TMEM lane quarter of this warp
        int ti = 0, ebox = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
            const int a = p.acc_stages == 2 ? (ti & 1) : 0;
            const uint32_t use = (uint32_t)(p.acc_stages == 2 ? (ti >> 1) : ti);
            mbar_wait(acc_full + a, use & 1, 5);
            tc_fence_after();
            const uint32_t tbase = tmem_base + (uint32_t)(a * acc_cols) + ((uint32_t)(q * 32) << 16);
            const long long row = (long long)t * kTileM + q * 32 + lane;
            const bool ok = row < p.R && !p.nostore;
            if (MODE == MODE_FWD) {
                float* orow = p.out + row * p.Fout;
                const bool sty = !POOL || p.out != nullptr;
                const bool wide8 = (p.Fout & 7) == 0 && (p.n0 & 7) == 0 && ((uintptr_t)p.out & 31) == 0;
                for (int c0 = 0; c0 < p.Nreal; c0 += 32) {
                    float v[32];
                    {
                        float lo[16];
                        tmem_ld16(tbase + c0, v);
                        tmem_ld16(tbase + p.Npad + c0, lo);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] += lo[i];
                    }
                    if (c0 + 16 < p.Npad) {
                        float lo[16];
                        tmem_ld16(tbase + c0 + 16, v + 16);
                        tmem_ld16(tbase + p.Npad + c0 + 16, lo);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[16 + i] += lo[i];
                    } else {
#pragma unroll
                        for (int i = 16; i < 32; ++i) v[i] = 0.f;
                    }
                    if (POOL) {
                        // HEALPix max-pool over 4 consecutive rows = 4 neighbouring lanes, on the raw y: batch norm without scale /
                        // shift, the bias and the activations are non-decreasing per feature, so the pooling commutes with them
                        // (models.py:1235-1243); the normalise / activate pass then reads a quarter of the rows
                        float m[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            m[i] = fmaxf(v[i], __shfl_xor_sync(0xffffffffu, v[i], 1));
                            m[i] = fmaxf(m[i], __shfl_xor_sync(0xffffffffu, m[i], 2));
                        }
                        if (ok && (lane & 3) == 0) {
                            float* prow = p.pool_out + (row >> 2) * p.Fout;
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {
                                const int c = p.n0 + c0 + i;
                                if (c + 3 < p.Fout && (p.Fout & 3) == 0) st4(prow + c, make_float4(m[i], m[i + 1], m[i + 2], m[i + 3]));
                                else {
#pragma unroll
                                    for (int e = 0; e < 4; ++e) if (c + e < p.Fout && c0 + i + e < p.Nreal) prow[c + e] = m[i + e];
                                }
                            }
                        }
                    }
                    if (!sty) {
                        // only the pooled output is wanted (inference)
                    } else if (ok && wide8) {
                        // a lane owns a whole output row: 32-byte stores halve the number of 128-byte lines (= load/store-pipe
                        // wavefronts) a warp instruction touches per byte (32 lines either way)
#pragma unroll
                        for (int i = 0; i < 32; i += 8) {
                            const int c = p.n0 + c0 + i;
                            if (c + 7 < p.Fout && c0 + i + 7 < p.Nreal) st8(orow + c, v + i);
                            else {
#pragma unroll
                                for (int e = 0; e < 8; ++e) if (c + e < p.Fout && c0 + i + e < p.Nreal) orow[c + e] = v[i + e];
                            }
                        }
                    } else if (ok) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const int c = p.n0 + c0 + i;
                            if (c + 3 < p.Fout && (p.Fout & 3) == 0) st4(orow + c, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
                            else {
#pragma unroll
                                for (int e = 0; e < 4; ++e) if (c + e < p.Fout && c0 + i + e < p.Nreal) orow[c + e] = v[i + e];
                            }
                        }
                    }
                    if (SUMS) {
                        // rows past R are zero (TMA zero fill), columns past Fout are zero weights
                        float w[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) w[i] = v[i] * v[i];
                        warp_column_sums(v, lane);
                        warp_column_sums(w, lane);
                        if (c0 + lane < p.Fout) { atomicAdd(red + c0 + lane, v[0]); atomicAdd(red + 256 + c0 + lane, w[0]); }
                    }
                }
            } else if (p.epi_tma) {
                // A lane owns a row: written straight to global memory, every store instruction of the warp touches 32
                // different 128-byte lines (26 / 33 us of the 99 / 83 us at cosmo layers 2 / 3 went into that).  The 128 rows of
                // one output plane of a tile are CONTIGUOUS in global memory (planes are [R][Fin] row-major), so the four
                // epilogue warps stage them row-major in shared memory and one lane stores the plane with a 1-D
                // cp.async.bulk (no load/store-pipe wavefronts for global memory; a TMA TENSOR store was tried first and is
                // limited to ~4 ns per box row).  A lane writes its four 16-byte chunks in an order rotated by its row index,
                // which spreads the lanes of a quarter warp over the bank lines (conflict-free at 64-byte rows, 2-way at
                // 128).  The warps share the issuing (a bulk copy costs its issuing warp ~0.3 us).  Per-plane fences and
                // barriers made the epilogue the bottleneck (120 us), so all K planes of the tile are staged first.
                const int bc = p.epi_cols;                              // = Fin = 16 (64-byte rows)
                const uint32_t buf_bytes = (uint32_t)kTileM * (uint32_t)bc * 4u;      // one plane of the tile
                const int ew = warp - kWarpEpi0, rloc = q * 32 + lane;
                const int rot = (rloc >> 1) & 3;
                const long long plane = p.R * (long long)p.Fin;
                const long long rows_left = p.R - (long long)t * kTileM;
                const uint32_t nbytes = (uint32_t)(rows_left < kTileM ? rows_left : kTileM) * (uint32_t)bc * 4u;
                // all planes of the tile are staged, then stored: one fence and two barriers per tile
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // my stores of the previous tile
                asm volatile("bar.sync 2, 128;" ::: "memory");
                for (int c0 = 0; c0 + 16 <= p.Nreal; c0 += 16) {
                    float hi[16], lo[16];
                    tmem_ld16(tbase + c0, hi);
                    tmem_ld16(tbase + p.Npad + c0, lo);
                    tmem_ld_wait();
                    const int kq = (p.n0 + c0) / bc;                    // Fin = 16: one round = one plane
                    float4 v0 = make_float4(hi[0] + lo[0], hi[1] + lo[1], hi[2] + lo[2], hi[3] + lo[3]);
                    float4 v1 = make_float4(hi[4] + lo[4], hi[5] + lo[5], hi[6] + lo[6], hi[7] + lo[7]);
                    float4 v2 = make_float4(hi[8] + lo[8], hi[9] + lo[9], hi[10] + lo[10], hi[11] + lo[11]);
                    float4 v3 = make_float4(hi[12] + lo[12], hi[13] + lo[13], hi[14] + lo[14], hi[15] + lo[15]);
                    if (rot & 1) { const float4 tmp = v0; v0 = v1; v1 = v2; v2 = v3; v3 = tmp; }
                    if (rot & 2) { float4 tmp = v0; v0 = v2; v2 = tmp; tmp = v1; v1 = v3; v3 = tmp; }
                    float4* rowp = reinterpret_cast<float4*>(epi + (size_t)kq * buf_bytes + (size_t)rloc * bc * 4);
                    rowp[(0 + rot) & 3] = v0;                           // instruction j carries chunk (j + rot) mod 4
                    rowp[(1 + rot) & 3] = v1;
                    rowp[(2 + rot) & 3] = v2;
                    rowp[(3 + rot) & 3] = v3;
                }
                fence_proxy_async();
                asm volatile("bar.sync 2, 128;" ::: "memory");
                if (lane == 0 && !p.nostore) {
                    for (int kq = ew; kq < p.K; kq += 4) {
                        float* dst = p.out + kq * plane + (long long)t * kTileM * p.Fin;
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                     :: "l"(dst), "r"(smem_u32(epi + (size_t)kq * buf_bytes)), "r"(nbytes) : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            } else {
                const long long plane = p.R * (long long)p.Fin;
                const bool wide = (((uintptr_t)p.out | (uintptr_t)(plane * 4)) & 31) == 0;   // 32-byte stores possible
                int c0 = 0;
                for (; c0 + 16 <= p.Nreal; c0 += 16) {              // 16 columns per TMEM round trip
                    float hi[16], lo[16];
                    tmem_ld16(tbase + c0, hi);
                    tmem_ld16(tbase + p.Npad + c0, lo);
                    tmem_ld_wait();
                    if (ok) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) hi[i] += lo[i];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int qq = p.n0 + c0 + 8 * h;
                            const int kq = qq / p.Fin, fi = qq - kq * p.Fin;
                            float* o = p.out + kq * plane + row * p.Fin + fi;
                            if (wide) st8(o, hi + 8 * h);
                            else { st4(o, make_float4(hi[8 * h], hi[8 * h + 1], hi[8 * h + 2], hi[8 * h + 3]));
                                   st4(o + 4, make_float4(hi[8 * h + 4], hi[8 * h + 5], hi[8 * h + 6], hi[8 * h + 7])); }
                        }
                    }
                }
                for (; c0 < p.Nreal; c0 += 8) {
                    float hi[8], lo[8];
                    tmem_ld8(tbase + c0, hi);
                    tmem_ld8(tbase + p.Npad + c0, lo);
                    tmem_ld_wait();
                    if (ok) {
                        const int qq = p.n0 + c0;
                        const int kq = qq / p.Fin, fi = qq - kq * p.Fin;
                        float* o = p.out + kq * plane + row * p.Fin + fi;
#pragma unroll
                        for (int i = 0; i < 8; ++i) hi[i] += lo[i];
                        if (wide) st8(o, hi);
                        else { st4(o, make_float4(hi[0], hi[1], hi[2], hi[3])); st4(o + 4, make_float4(hi[4], hi[5], hi[6], hi[7])); }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(acc_empty + a);
        }
        if (MODE == MODE_BWD_DATA && p.epi_tma && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        (void)ebox;
        if (SUMS) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int e = threadIdx.x - kWarpEpi0 * 32; e < p.Fout; e += 128) {
                atomicAdd(p.bn_sums + e, (double)red[e]);
                atomicAdd(p.bn_sums + p.Fout + e, (double)red[256 + e]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kWarpMma) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

struct RowGemmPlan {
    int kc, swz, nchunk, nkb, Npad, G, nstages, acc_stages, wstream, tmem_cols, ctas;
    size_t smem, wbytes;
};

#define env_int(name, dflt) DSB_ENV_INT(name, dflt)

// Shared-memory plan; returns false if the shape does not fit the tensor-core path.
static bool plan_rowgemm_n(int red_per_plane, int planes, int Npad, bool can_stream, int ctas, RowGemmPlan* pl, size_t reserve);

// Try `ctas` CTAs per SM first (independent pipelines hide the TMA -> split -> MMA -> epilogue latency chain of a
// CTA), fall back to one.
// `reserve`: bytes kept free after the pipeline's own layout (epilogue staging of the data gradient)
static bool plan_rowgemm(int red_per_plane, int planes, int Npad, bool can_stream, int ctas, RowGemmPlan* pl, size_t reserve = 0)
{
    for (; ctas >= 1; --ctas)
        if (plan_rowgemm_n(red_per_plane, planes, Npad, can_stream && ctas == 1, ctas, pl, reserve) && (ctas == 1 || pl->nstages >= 3)) return true;
    return false;
}

static bool plan_rowgemm_n(int red_per_plane, int planes, int Npad, bool can_stream, int ctas, RowGemmPlan* pl, size_t reserve)
{
    if (red_per_plane % 8) return false;
    pl->kc = red_per_plane % 32 == 0 ? 32 : red_per_plane % 16 == 0 ? 16 : 8;
    pl->swz = pl->kc == 32 ? 3 : pl->kc == 16 ? 2 : 1;
    pl->nchunk = red_per_plane / pl->kc;
    pl->nkb = planes * pl->nchunk;
    pl->Npad = Npad;
    pl->G = 32 / pl->kc;                                   // 16 KB of activations per stage
    if (pl->G > pl->nkb) pl->G = pl->nkb;
    pl->wbytes = (size_t)pl->nkb * 2 * Npad * pl->kc * 4;
    const size_t fixed = 1024 /*align*/ + 3 * kMaxStages * 8 + 4 * 8 + 16 + 2 * 256 * 4 + 64 + reserve;
    const size_t budget = (size_t)(227 * 1024) / ctas - (ctas > 1 ? 1024 : 0);
    const size_t min_smem = (size_t)(227 * 1024) / (ctas + 1) + 1024;     // never more CTAs per SM than planned (TMEM)
    pl->ctas = ctas;
    pl->acc_stages = (4 * Npad * ctas <= 512) ? 2 : 1;
    if (pl->acc_stages * 2 * Npad * ctas > 512) return false;
    pl->tmem_cols = 32;
    while (pl->tmem_cols < pl->acc_stages * 2 * Npad) pl->tmem_cols <<= 1;
    if (pl->tmem_cols * ctas > 512) return false;
    size_t stage = 2 * (size_t)pl->G * kTileM * pl->kc * 4;
    pl->wstream = 0;
    if (can_stream && pl->wbytes + fixed + 4 * stage > budget) {
        // resident weights would leave fewer than 4 stages: stream the weight blocks with the activations
        pl->wstream = 1;
        pl->G = 1;
        stage = 2 * (size_t)kTileM * pl->kc * 4 + 2 * (size_t)Npad * pl->kc * 4;
        if (fixed + 2 * stage > budget) return false;
        size_t ns = (budget - fixed) / stage;
        if (ns > kMaxStages) ns = kMaxStages;
        pl->nstages = (int)ns;
        pl->smem = fixed + ns * stage;
        if (pl->smem < min_smem) pl->smem = min_smem;
        return true;
    }
    while (pl->G > 1 && pl->wbytes + fixed + 3 * stage > budget) { pl->G >>= 1; stage >>= 1; }
    if (pl->wbytes + fixed + 2 * stage > budget) return false;
    size_t ns = (budget - pl->wbytes - fixed) / stage;
    if (ns > kMaxStages) ns = kMaxStages;
    pl->nstages = (int)ns;
    pl->smem = pl->wbytes + fixed + ns * stage;
    if (pl->smem < min_smem) pl->smem = min_smem;
    return true;
}

// CTAs per SM the register file allows for this instance (65536 registers per SM)
template <int MODE, bool SUMS>
static int rowgemm_max_ctas()
{
    static int cached = 0;
    if (!cached) {
        cudaFuncAttributes a;
        cached = 1;
        if (cudaFuncGetAttributes(&a, rowgemm_tc_kernel<MODE, SUMS>) == cudaSuccess && a.numRegs > 0) {
            const int regs = (a.numRegs + 7) / 8 * 8;
            cached = 65536 / (regs * kRgThreads);
            if (cached < 1) cached = 1;
        }
    }
    return cached;
}

template <int MODE, bool SUMS, bool POOL = false>
static int launch_rowgemm(RowGemmParams& p, size_t smem, int ctas, float* image, cudaStream_t s)
{
    static bool configured = false;                // per template instance
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(rowgemm_tc_kernel<MODE, SUMS, POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    p.wpack = image;
    p.nostore = DSB_EXPERIMENT_INT("DSB200_RG_NOSTORE", 0);
#ifdef DSB200_EXPERIMENTS
    { const char* d = getenv("DSB200_RG_DBG"); p.dbg = (d && *d) ? (long long*)strtoull(d, nullptr, 0) : nullptr; }
#else
    p.dbg = nullptr;
#endif
    if (image) {
        const int total = p.nkb * 2 * p.Npad * p.kc;
        pack_w_kernel<MODE><<<(total + 255) / 256, 256, 0, s>>>(p, image);
        count_launch();
    }
    const int slots = kNumSMs * ctas;
    const int grid = p.ntiles < slots ? p.ntiles : slots;
    rowgemm_tc_kernel<MODE, SUMS, POOL><<<grid, kRgThreads, smem, s>>>(p);
    DSB_LAUNCH_CHECK();
}

size_t contract_workspace_bytes(int Fin, int K, int Fout)
{
    const size_t Q = (size_t)K * Fin;
    const size_t fwd = 2 * (size_t)((Fout + 15) / 16 * 16) * Q * 4;
    const size_t nb = (Q + 159) / 160;
    const size_t NB = ((Q + nb - 1) / nb + 15) / 16 * 16;
    const size_t bwd = 2 * nb * NB * (size_t)Fout * 4;
    size_t need = (fwd > bwd ? fwd : bwd) + 1024;
    const size_t cols = colgemm_workspace_bytes(Fin, K);
    return need > cols ? need : cols;
}

// Forward contraction on the tensor cores.  Returns 1 if the shape is not supported (caller falls back).
int contract_fwd_tc(const float* x0, const float* xk, const float* W, float* Y, long long R, int Fin, int K, int Fout,
                    double* bn_sums, void* workspace, cudaStream_t s, float* Ypool)
{
    if (R < kTileM || R > 0x7fffff00LL || Fout > 256) return 1;
    if ((((uintptr_t)x0 | (uintptr_t)xk | (uintptr_t)Y | (uintptr_t)workspace | (uintptr_t)Ypool) & 15) != 0) return 1;
    if (Ypool && (R & 3)) return 1;
    const int Npad = (Fout + 15) / 16 * 16;
    RowGemmPlan pl;
    int ctas = env_int("DSB200_RG_CTAS", 2);
    const int rmax = bn_sums ? rowgemm_max_ctas<MODE_FWD, true>() : rowgemm_max_ctas<MODE_FWD, false>();
    if (ctas > rmax) ctas = rmax;
    if (!plan_rowgemm(Fin, K, Npad, workspace != nullptr, ctas, &pl)) return 1;
    RowGemmParams p;
    p.tmem_cols = pl.tmem_cols;
    if (make_tmap_f32(&p.tmA0, x0, Fin, R, 1, pl.kc, kTileM, pl.swz, false)) return 1;
    if (K > 1) { if (make_tmap_f32(&p.tmAk, xk, Fin, R, K - 1, pl.kc, kTileM, pl.swz, true)) return 1; }
    else p.tmAk = p.tmA0;
    p.W = W; p.out = Y; p.pool_out = Ypool; p.bn_sums = bn_sums; p.R = R; p.Fin = Fin; p.K = K; p.Fout = Fout;
    p.epi_tma = 0; p.epi_cols = 0; p.tmOut = p.tmA0;
    p.kc = pl.kc; p.nchunk = pl.nchunk; p.nkb = pl.nkb; p.G = pl.G; p.Npad = Npad; p.Nreal = Fout; p.n0 = 0; p.swz = pl.swz;
    p.nstages = pl.nstages; p.acc_stages = pl.acc_stages; p.wstream = pl.wstream;
    p.ntiles = (int)((R + kTileM - 1) / kTileM);
    float* image = (float*)workspace;
    if (Ypool) {
        if (!bn_sums) return launch_rowgemm<MODE_FWD, false, true>(p, pl.smem, pl.ctas, image, s);
        return launch_rowgemm<MODE_FWD, true, true>(p, pl.smem, pl.ctas, image, s);
    }
    if (!bn_sums) return launch_rowgemm<MODE_FWD, false>(p, pl.smem, pl.ctas, image, s);
    return launch_rowgemm<MODE_FWD, true>(p, pl.smem, pl.ctas, image, s);
}

// 1 when contract_fwd_tc takes the shape (alignment aside)
int contract_fwd_tc_query(long long R, int Fin, int K, int Fout)
{
    if (R < kTileM || R > 0x7fffff00LL || Fout > 256) return 0;
    RowGemmPlan pl;
    int ctas = env_int("DSB200_RG_CTAS", 2);
    const int rmax = rowgemm_max_ctas<MODE_FWD, true>();
    if (ctas > rmax) ctas = rmax;
    return plan_rowgemm(Fin, K, (Fout + 15) / 16 * 16, true, ctas, &pl) ? 1 : 0;
}

// Data gradient on the tensor cores: G[k][r][fi] = sum_fo dY[r, fo] W[fi*K + k, fo].
int contract_bwd_data_tc(const float* dY, const float* W, float* G, long long R, int Fin, int K, int Fout,
                         void* workspace, cudaStream_t s)
{
    if (R < kTileM || R > 0x7fffff00LL || (Fin & 7) || (Fout & 7)) return 1;
    if ((((uintptr_t)dY | (uintptr_t)G | (uintptr_t)workspace) & 15) != 0) return 1;
    const int Q = K * Fin;
    // n-blocks of at most 160 output columns (hi + lo accumulators = 320 TMEM columns)
    const int nb = (Q + 159) / 160;
    const int NB = ((Q + nb - 1) / nb + 15) / 16 * 16;
    RowGemmPlan pl;
    int ctas = env_int("DSB200_RG_CTAS", 2);
    if (ctas > rowgemm_max_ctas<MODE_BWD_DATA, false>()) ctas = rowgemm_max_ctas<MODE_BWD_DATA, false>();
    // bulk-store epilogue (experiment, off by default: DSB200_RG_BULK_STORE=1): planes of 16 columns (cosmo layer 2), all K
    // planes of a 128-row tile staged at once.  Measured at layer 2: 117 us against 99 us with direct 32-byte stores -- the
    // staging alone costs 10 us and the single buffer makes every tile wait for the previous tile's stores to be read.
    const int epi_cols = (Fin == 16 && nb == 1 && K <= 6 && (((uintptr_t)G) & 15) == 0 && env_int("DSB200_RG_BULK_STORE", 0)) ? Fin : 0;
    const size_t reserve = epi_cols ? 1024 + (size_t)K * kTileM * epi_cols * 4 : 0;
    bool epi_tma = epi_cols != 0;
    if (!epi_tma || !plan_rowgemm(Fout, 1, NB, workspace != nullptr, ctas, &pl, reserve)) {
        epi_tma = false;
        if (!plan_rowgemm(Fout, 1, NB, workspace != nullptr, ctas, &pl)) return 1;
    }
    for (int b = 0; b < nb; ++b) {
        RowGemmParams p;
        p.tmem_cols = pl.tmem_cols;
        if (make_tmap_f32(&p.tmA0, dY, Fout, R, 1, pl.kc, kTileM, pl.swz, false)) return 1;
        p.tmAk = p.tmA0;
        p.epi_tma = epi_tma ? 1 : 0; p.epi_cols = epi_cols; p.tmOut = p.tmA0;
        p.W = W; p.out = G; p.pool_out = nullptr; p.bn_sums = nullptr; p.R = R; p.Fin = Fin; p.K = K; p.Fout = Fout;
        p.kc = pl.kc; p.nchunk = pl.nchunk; p.nkb = pl.nkb; p.G = pl.G; p.Npad = NB; p.n0 = b * NB;
        p.Nreal = (Q - p.n0 < NB) ? (Q - p.n0) : NB;
        p.swz = pl.swz; p.nstages = pl.nstages; p.acc_stages = pl.acc_stages; p.wstream = pl.wstream;
        p.ntiles = (int)((R + kTileM - 1) / kTileM);
        float* image = workspace ? (float*)((uint8_t*)workspace + (size_t)b * pl.wbytes) : nullptr;
        int rc = launch_rowgemm<MODE_BWD_DATA, false>(p, pl.smem, pl.ctas, image, s);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace tc
}  // namespace dsb200

// =====================================================================================================
// Weight gradient on the tensor cores:  dW[fi*K + k, fo] += sum_r X_k[r, fi] * dY[r, fo]   (autodiff of
// tf.matmul, models.py:1077).  The reduction runs over the (sample, vertex) rows, so both operands are
// "MN-major" in UMMA terms.  For 32-bit (TF32) MN-major operands the only shared-memory layout the tensor
// core accepts is the 128-byte swizzle with 32-byte atomicity: a TMA box of RS rows x 32 columns
// (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; narrower planes are zero-filled to 32 columns) lands as that
// canonical layout, one box per 32-column block of every plane of X and of dY.
//   D[q, fo] (TMEM, one 128-lane tile per 128 values of q)  +=  Xh^T [dYh | dYl]  +  Xl^T dYh
// Every CTA owns a contiguous range of rows, accumulates in TMEM over its whole range and adds its
// partial result to dW with one fp32 atomic per entry at the end.
namespace dsb200 {
namespace tc {

constexpr int kMNBlock = 32;                     // columns per MN-major block (128 bytes of fp32)
constexpr int kWgThreads = 448;                  // warp 0 TMA, warp 1 MMA, warps 2-5 / 6-9 split (alternate stages), warps 10-13 epilogue

struct alignas(64) WGradParams {
    CUtensorMap tmX0, tmXk, tmdY;
    float* dW;
    long long R;
    int Fin, K, Fout;
    int nblkA, nblkB;          // 32-column blocks per plane of X / of dY (after row packing)
    int pack;                  // unused plane packing experiment (always 0)
    int rpk;                   // Fin < 32: rpk = 32/Fin consecutive rows are viewed as ONE row of rpk*Fin features (and of
                               // rpk*Fout gradient columns); only the diagonal blocks (same sub-row) of D are used
    int tmem_cols;             // TMEM columns to allocate (power of two >= mtiles * 2 * Np)
    int vec_red;               // dW rows are 16-byte aligned: red.global.add.v4.f32
    int swap;                  // 1: dY on the M side ([dYh; dYl] stacked, 128 lanes), X blocks on the N side (see below)
    int nNb, bpn;              // swap: number of N blocks, X blocks per N block (N = 32 * bpn <= 256)
    int nA;                    // number of X blocks per stage
    int pad_blocks;            // unused blocks at the end of a stage (the last 128-wide tile of q may read past X_lo)
    int RS;                    // rows per pipeline stage (multiple of 8)
    int nstages;
    int mtiles;                // ceil(K * nblkA * 32 / 128)
    long long iters_total;     // ceil(R / RS)
    long long iters_per_cta;
};

// descriptor of an MN-major TF32 operand: layout type 1 (128B swizzle, 32B atoms), 4-row k-groups
__device__ __forceinline__ uint64_t make_mn_desc(uint32_t saddr, uint32_t block_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((block_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((512u >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (1ull << 61);
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ WGradParams p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.nstages;
    const uint32_t blk = (uint32_t)p.RS * kMNBlock * 4;               // one 32-column block of RS rows
    const int nA = p.nA;
    const uint32_t xbytes = blk * nA, ybytes = blk * p.nblkB;
    const uint32_t stage_bytes = 2 * (xbytes + ybytes) + blk * p.pad_blocks;   // [X][X_lo][dY][dY_lo][pad]
    const int Np = p.nblkB * kMNBlock;                                // dY columns incl. zero padding

    uint8_t* sA = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)S * stage_bytes);
    uint64_t* full_tma = bars;
    uint64_t* full_mma = bars + kMaxStages;
    uint64_t* empty = bars + 2 * kMaxStages;
    uint64_t* acc_full = bars + 3 * kMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full_tma + s, 1); mbar_init(full_mma + s, 128); mbar_init(empty + s, 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
        prefetch_tmap(&p.tmX0);
        prefetch_tmap(&p.tmXk);
        prefetch_tmap(&p.tmdY);
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const long long it0 = (long long)blockIdx.x * p.iters_per_cta;
    long long it1 = it0 + p.iters_per_cta;
    if (it1 > p.iters_total) it1 = p.iters_total;
    const int niter = it1 > it0 ? (int)(it1 - it0) : 0;
    const int acc_cols = 2 * Np;                                      // per M-tile: [hi | lo]

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < niter; ++i) {
                const int row0 = (int)((it0 + i) * p.RS);
                mbar_wait(empty + s, ph ^ 1u, 11);
                mbar_expect_tx(full_tma + s, xbytes + ybytes);
                uint8_t* dst = sA + (size_t)s * stage_bytes;
                if (p.pack > 0) {
                    // block 0 = plane 0 (zero-filled to 32 columns); block b >= 1 = planes 1+(b-1)*pack .. of the basis
                    tma_load_2d(dst, &p.tmX0, full_tma + s, 0, row0);
                    for (int bq = 1; bq < nA; ++bq)
                        tma_load_3d(dst + (size_t)bq * blk, &p.tmXk, full_tma + s, 0, (bq - 1) * p.pack, row0);
                } else {
                    for (int k = 0; k < p.K; ++k)
                        for (int c = 0; c < p.nblkA; ++c) {
                            uint8_t* d = dst + (size_t)(k * p.nblkA + c) * blk;
                            if (k == 0) tma_load_2d(d, &p.tmX0, full_tma + s, c * kMNBlock, row0);
                            else tma_load_3d(d, &p.tmXk, full_tma + s, c * kMNBlock, row0, k - 1);
                        }
                }
                uint8_t* dy = dst + 2 * (size_t)xbytes;
                for (int c = 0; c < p.nblkB; ++c) tma_load_2d(dy + (size_t)c * blk, &p.tmdY, full_tma + s, c * kMNBlock, row0);
                if (++s == S) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && niter > 0) {
            const bool merged = acc_cols <= 256;
            const uint32_t idesc_wide = make_idesc_tf32(kTileM, merged ? acc_cols : Np, 1, 1);
            const uint32_t idesc_half = make_idesc_tf32(kTileM, Np, 1, 1);
            const uint64_t dA0 = make_mn_desc(smem_u32(sA), blk);
            const uint64_t dB0 = make_mn_desc(smem_u32(sA) + 2 * xbytes, blk);
            const uint32_t kstep = (8u * kMNBlock * 4) >> 4;                         // 8 rows of 128 bytes
            const uint32_t mtA = ((uint32_t)(kTileM / kMNBlock) * blk) >> 4;         // next 128 values of q
            const uint32_t loA = xbytes >> 4, loB = ybytes >> 4, stg = stage_bytes >> 4;
            const int ksteps = p.RS / 8;
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < niter; ++i) {
                mbar_wait(full_mma + s, ph, 13);
                tc_fence_after();
                const uint64_t dA = dA0 + (uint64_t)(stg * (uint32_t)s), dB = dB0 + (uint64_t)(stg * (uint32_t)s);
                if (p.swap) {
                    // One tcgen05.mma costs ~87 ns whatever its M and N (tools/experiments/mma_rate.cu), so the widest
                    // operand goes on N: D[(hi|lo, fo), q] += [dYh; dYl]^T (X_h + X_l), two instructions per 8 rows
                    // and N block (the lo x lo term it adds is below fp32 resolution).
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint32_t acc = (i > 0 || ks > 0) ? 1u : 0u;
                        const uint64_t a = dB + kstep * ks;                           // [dYh | dYl | pad] blocks: M = 128
                        for (int nb = 0; nb < p.nNb; ++nb) {
                            const int nblk = min(p.bpn, nA - nb * p.bpn);
                            const uint32_t idesc = make_idesc_tf32(kTileM, nblk * kMNBlock, 1, 1);
                            const uint64_t b = dA + kstep * ks + (uint64_t)(((uint32_t)(nb * p.bpn) * blk) >> 4);
                            const uint32_t d = tmem_base + (uint32_t)(nb * p.bpn * kMNBlock);
                            mma_tf32(d, a, b, idesc, acc);
                            mma_tf32(d, a, b + loA, idesc, 1u);
                        }
                    }
                    tc_commit(empty + s);
                    if (++s == S) { s = 0; ph ^= 1u; }
                    continue;
                }
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint32_t acc = (i > 0 || ks > 0) ? 1u : 0u;
                    const uint64_t b = dB + kstep * ks;
                    for (int mt = 0; mt < p.mtiles; ++mt) {
                        const uint64_t a = dA + kstep * ks + mtA * mt;
                        const uint32_t d = tmem_base + (uint32_t)(mt * acc_cols);
                        if (merged) {
                            mma_tf32(d, a, b, idesc_wide, acc);                   // Xh^T [dYh | dYl]
                        } else {
                            mma_tf32(d, a, b, idesc_half, acc);
                            mma_tf32(d + Np, a, b + loB, idesc_half, acc);
                        }
                        mma_tf32(d, a + loA, b, idesc_half, 1u);                  // Xl^T dYh
                    }
                }
                tc_commit(empty + s);
                if (++s == S) { s = 0; ph ^= 1u; }
            }
            tc_commit(acc_full);
        }
    } else if (warp < kWarpEpi0) {
        const int grp = (warp - kWarpSplit0) >> 2;
        const int tt = (threadIdx.x - kWarpSplit0 * 32) & 127;
        const int nvecX = (int)(xbytes / 16), nvecY = (int)(ybytes / 16);
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < niter; ++i) {
            if ((s & 1) != grp) { if (++s == S) { s = 0; ph ^= 1u; } continue; }     // stage -> group, see the row-GEMM
            mbar_wait(full_tma + s, ph, 14);
            float4* X = reinterpret_cast<float4*>(sA + (size_t)s * stage_bytes);
            float4* Xl = X + nvecX;
            float4* Y = Xl + nvecX;
            float4* Yl = Y + nvecY;
#pragma unroll 4
            for (int j = tt; j < nvecX + nvecY; j += 128) {
                float4* src = j < nvecX ? X + j : Y + (j - nvecX);
                float4* dlo = j < nvecX ? Xl + j : Yl + (j - nvecX);
                *dlo = tf32_dropped_bits(*src);                   // hi = the tile as it landed (read truncated by the tensor core)
            }
            fence_proxy_async();
            mbar_arrive(full_mma + s);
            if (++s == S) { s = 0; ph ^= 1u; }
        }
    } else if (niter > 0) {
        // epilogue: once per CTA
        const int qd = warp & 3;
        mbar_wait(acc_full, 0, 15);
        tc_fence_after();
        const int Fx = p.Fin * p.rpk;                                // features of a packed row of X
        if (p.swap) {
            // lanes: (hi | lo part, gradient column), columns: q.  Lanes of a warp hold consecutive fo: coalesced reductions.
            const int Fy = p.Fout * p.rpk;
            const int L = qd * 32 + lane;
            const int part = L / Np, cy = L - part * Np;             // Np = gradient columns incl. the zero fill of the last block
            const int hy = cy / p.Fout, fo = cy - hy * p.Fout;
            const bool lane_ok = part < 2 && cy < Fy;
            const uint32_t tbase = tmem_base + ((uint32_t)(qd * 32) << 16);
            const int ncols = nA * kMNBlock;
            for (int c0 = 0; c0 < ncols; c0 += 16) {
                float v[16];
                tmem_ld16(tbase + c0, v);
                tmem_ld_wait();
                const int bq = c0 / kMNBlock;
                const int k = bq / p.nblkA;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int cc = (bq - k * p.nblkA) * kMNBlock + (c0 % kMNBlock) + e;   // column within the packed plane
                    const int hx = cc / p.Fin, fi = cc - hx * p.Fin;
                    if (lane_ok && hx == hy && cc < Fx) atomicAdd(p.dW + ((long long)fi * p.K + k) * p.Fout + fo, v[e]);
                }
            }
            tc_fence_before();
        } else
        for (int mt = 0; mt < p.mtiles; ++mt) {
            const int q = mt * kTileM + qd * 32 + lane;              // (block, column in block) of the padded Xcat
            const int bq = q / kMNBlock, cq = q - bq * kMNBlock;
            const int k = bq / p.nblkA;
            const int cc = (bq - k * p.nblkA) * kMNBlock + cq;       // column within the packed plane
            const int h = cc / p.Fin, fi = cc - h * p.Fin;           // sub-row, feature
            const bool ok = bq < nA && k < p.K && cc < Fx;
            float* dst = p.dW + ((long long)fi * p.K + k) * p.Fout;
            const uint32_t tbase = tmem_base + (uint32_t)(mt * acc_cols) + ((uint32_t)(qd * 32) << 16);
            for (int c0 = 0; c0 < p.Fout; c0 += 8) {
                for (int hh = 0; hh < p.rpk; ++hh) {                 // TMEM loads are warp-uniform: read every sub-row block
                    float hi[8], lo[8];
                    tmem_ld8(tbase + hh * p.Fout + c0, hi);
                    tmem_ld8(tbase + Np + hh * p.Fout + c0, lo);
                    tmem_ld_wait();
                    if (ok && hh == h) {
                        if (p.vec_red) {
                            red_add_v4(dst + c0, hi[0] + lo[0], hi[1] + lo[1], hi[2] + lo[2], hi[3] + lo[3]);
                            red_add_v4(dst + c0 + 4, hi[4] + lo[4], hi[5] + lo[5], hi[6] + lo[6], hi[7] + lo[7]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e) atomicAdd(dst + c0 + e, hi[e] + lo[e]);
                        }
                    }
                }
            }
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

int contract_bwd_weight_tc(const float* x0, const float* xk, const float* dY, float* dW, long long R, int Fin, int K,
                           int Fout, cudaStream_t s)
{
    if (R < 128 || R > 0x7fffff00LL || (Fin & 7) || (Fout & 7) || Fout > 256) return 1;
    if ((((uintptr_t)x0 | (uintptr_t)xk | (uintptr_t)dY) & 15) != 0) return 1;
    WGradParams p;
    // Narrow planes (Fin < 32): view rpk = 32/Fin consecutive rows as one row, X [R/rpk, 32] and dY [R/rpk, rpk*Fout].
    // Every TMA row is then a full 128-byte line and no shared memory holds zero padding; the product has rpk x rpk
    // blocks of which the rpk diagonal ones add up to dW (the tensor pipe has the headroom for the rest).
    int rpk = 1;
    if (Fin < kMNBlock && kMNBlock % Fin == 0 && R % (kMNBlock / Fin) == 0 && (kMNBlock / Fin) * Fout <= 256) rpk = kMNBlock / Fin;
    if (env_int("DSB200_WG_NOPACK", 0)) rpk = 1;
    p.rpk = rpk;
    const long long Rp = R / rpk;
    const int Fx = Fin * rpk, Fy = Fout * rpk;
    p.nblkA = (Fx + kMNBlock - 1) / kMNBlock;
    p.nblkB = (Fy + kMNBlock - 1) / kMNBlock;
    p.pack = 0;
    const int nA = K * p.nblkA;
    p.nA = nA;
    const int nB = p.nblkB;
    p.mtiles = (nA * kMNBlock + kTileM - 1) / kTileM;
    // swapped roles when [dYh; dYl] fits the 128 lanes of one accumulator
    p.swap = (nB <= 2 && env_int("DSB200_WG_SWAP", 0)) ? 1 : 0;     // measured: no faster (the kernel is latency bound, not MMA bound)
    p.nNb = (nA * kMNBlock + 255) / 256;
    p.bpn = (nA + p.nNb - 1) / p.nNb;
    const int need_cols = p.swap ? nA * kMNBlock : p.mtiles * 2 * nB * kMNBlock;
    if (need_cols > 512) return 1;
    int tcols = 32;
    while (tcols < need_cols) tcols <<= 1;
    p.tmem_cols = tcols;
    // CTAs per SM: limited by TMEM (512 columns) and by shared memory
    int ctas = 512 / tcols;
    const int want = env_int("DSB200_WG_CTAS", 2);
    if (ctas > want) ctas = want;
    if (ctas < 1) ctas = 1;
    // the last 128-wide tile of q may read past the X_lo blocks: keep it inside the stage ([X][X_lo][dY][dY_lo][pad])
    p.pad_blocks = p.mtiles * (kTileM / kMNBlock) - (nA + 2 * nB);
    if (p.swap) p.pad_blocks = kTileM / kMNBlock - 2 * nB;         // the 128-lane operand [dYh | dYl | pad] stays inside the stage
    if (p.pad_blocks < 0) p.pad_blocks = 0;
    const int per_row = kMNBlock * 4 * (2 * (nA + nB) + p.pad_blocks);
    const int stage_kb = env_int("DSB200_WG_STAGE_KB", ctas > 1 ? 32 : 48);
    int RS = (stage_kb * 1024) / per_row / 8 * 8;
    if (RS > 128) RS = 128;
    if (RS < 8) RS = 8;
    if (RS > Rp) return 1;
    p.RS = RS;
    const size_t stage = (size_t)RS * per_row;
    const size_t fixed = 1024 + 3 * kMaxStages * 8 + 64;
    const size_t budget = (size_t)(227 * 1024) / ctas - (ctas > 1 ? 1024 : 0);      // 1 KB per CTA is reserved by the driver
    if (budget < fixed + 2 * stage) return 1;
    size_t ns = (budget - fixed) / stage;
    if (ns > kMaxStages) ns = kMaxStages;
    p.nstages = (int)ns;
    size_t smem = fixed + ns * stage;
    const size_t min_smem = (size_t)(227 * 1024) / (ctas + 1) + 1024;                  // never more CTAs per SM than TMEM allows
    if (smem < min_smem) smem = min_smem;
    const int SW = 4;                                                   // 128B swizzle, 32B atoms
    if (make_tmap_f32(&p.tmX0, x0, Fx, Rp, 1, kMNBlock, RS, SW, false)) return 1;
    if (K > 1) { if (make_tmap_f32(&p.tmXk, xk, Fx, Rp, K - 1, kMNBlock, RS, SW, true)) return 1; }
    else p.tmXk = p.tmX0;
    if (make_tmap_f32(&p.tmdY, dY, Fy, Rp, 1, kMNBlock, RS, SW, false)) return 1;
    p.dW = dW; p.R = Rp; p.Fin = Fin; p.K = K; p.Fout = Fout;
    p.vec_red = (((uintptr_t)dW & 15) == 0) ? 1 : 0;
    p.iters_total = (Rp + RS - 1) / RS;
    const long long slots = (long long)kNumSMs * ctas;
    int grid = p.iters_total < slots ? (int)p.iters_total : (int)slots;
    p.iters_per_cta = (p.iters_total + grid - 1) / grid;
    grid = (int)((p.iters_total + p.iters_per_cta - 1) / p.iters_per_cta);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    wgrad_tc_kernel<<<grid, kWgThreads, smem, s>>>(p);
    DSB_LAUNCH_CHECK();
}

}  // namespace tc
}  // namespace dsb200
