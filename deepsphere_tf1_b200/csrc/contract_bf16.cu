// K2 in the BF16 mode: the dense contraction of cgcnn.chebyshev5 on bfloat16 activations,
//   Y[r, fo]     = sum_{k, fi} X_k[r, fi] * W[fi*K + k, fo]          (reference deepsphere/models.py:1062-1078)
//   dX_k[r, fi]  = sum_fo dY[r, fo] * W[fi*K + k, fo]                (autodiff of models.py:1077)
//   dW[fi*K+k,fo] += sum_r X_k[r, fi] * dY[r, fo]
// as tcgen05.mma kind::f16 (BF16 x BF16 -> FP32 in TMEM).  What the fp32 kernels (contract_tc.cu) spend their time on -- the
// TF32 hi / lo split of every landed tile through shared memory -- does not exist here: a TMA tensor copy lands a tile in the
// canonical swizzled layout and the tensor core reads it as it is.  Design points, all from the round-1 measurements:
//   * TMA tensor copies reach the HBM rate only with >= 16 KB per box (profiles/r01_tma_box_rate.txt): the K - 1 planes of the
//     Chebyshev basis of a row tile come as ONE rank-3 box (columns, rows, planes), and tiles of 64-byte rows are 256 rows;
//   * narrow rows are never padded: rp consecutive rows are viewed as one row of rp*F elements (the planes are row-major, so
//     this is a pure re-interpretation) and the weights become block diagonal -- the tensor pipe has an order of magnitude of
//     headroom at these arithmetic intensities, shared memory and TMA do not;
//   * persistent CTAs, one per SM: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 / 6-9 = two epilogue groups that take
//     alternate tiles (double-buffered TMEM accumulators), weights resident in shared memory as a pre-packed bf16 image.
#include "common.cuh"
#include "bf16.cuh"
#include "tc_common.cuh"

namespace dsb200 {
namespace tc {

// bf16 tensor map: dims / box innermost first, strides (bytes) for dims 1..rank-1; swz 2 / 3 = 64 / 128-byte swizzle
static int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                          const cuuint32_t* box, int swz)
{
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return -1;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUtensorMapSwizzle sw = swz == 3 ? CU_TENSOR_MAP_SWIZZLE_128B : swz == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swz == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

// D[tmem] (+)= A[smem] * B[smem], BF16 inputs, FP32 accumulate
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// instruction descriptor, kind::f16: c_format F32 (1) at [4,6), a_format / b_format BF16 (1) at [7,10) / [10,13)
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void st_b32x8(void* dst, const uint32_t* w) {      // one 32-byte store, dst 32-byte aligned
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "l"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}

constexpr int kBM = 128;                 // rows of one MMA (TMEM lanes)
constexpr int kBfThreads = 320;          // warp 0 TMA, warp 1 MMA, warps 2-5 / 6-9 epilogue groups
constexpr int kBfMaxStages = 6;
enum { BF_FWD = 0, BF_BWD_DATA = 1 };

struct alignas(64) RowGemmBfParams {
    CUtensorMap tmA0;          // plane 0 of the (row-packed) activations, or dY
    CUtensorMap tmAk;          // planes 1..K-1, rank 3 (forward only)
    const float* W;            // [Fin*K, Fout], row = fi*K + k
    const uint8_t* wimage;     // packed bf16 weights in the shared-memory image layout (pack_w_bf16_kernel)
    bf16* out;                 // Y [R, Fout] (forward: may be NULL when only the pooled output is wanted)  |  G [K][R][Fin]
    bf16* pool_out;            // forward only, may be NULL: max over every 4 consecutive rows of Y, [R/4, Fout]
    double* bn_sums;           // forward only, may be NULL
    long long Rp;              // packed rows = R / rp
    int Fin, K, Fout, rp;
    int rowbytes, kc;          // bytes / elements of one k-block row (64 / 32 or 128 / 64)
    int nkb, nchunk;           // k-blocks per tile; forward: chunks per plane (k-block = chunk * K + plane)
    int Ntot, Nsub, nsplit;    // accumulator columns = nsplit MMAs of Nsub columns
    int Nreal;                 // valid output columns of a packed row
    int swz, TM;               // swizzle mode; 128-row MMA tiles per stage
    int nstages, acc_stages, tmem_cols, ntiles;
};

// Element (j, n, kk) of the weight image: k-block j, output column n, reduction element kk of the block.
template <int MODE>
__device__ __forceinline__ float w_bf_value(const RowGemmBfParams& p, int j, int n, int kk)
{
    if (n >= p.Nreal) return 0.f;
    if (MODE == BF_FWD) {
        const int c = j / p.K, k = j - c * p.K;
        const int pc = c * p.kc + kk;                      // column of the packed input row: (sub-row h, feature fi)
        const int h = pc / p.Fin, fi = pc - h * p.Fin;
        const int h2 = n / p.Fout, fo = n - h2 * p.Fout;   // column of the packed output row
        return h == h2 ? __ldg(p.W + ((long long)fi * p.K + k) * p.Fout + fo) : 0.f;
    } else {
        const int pc = j * p.kc + kk;                      // column of the packed dY row: (h, fo)
        const int h = pc / p.Fout, fo = pc - h * p.Fout;
        const int per = p.rp * p.Fin;                      // output columns are ordered (plane k, sub-row h2, feature fi)
        const int k = n / per, rem = n - k * per;
        const int h2 = rem / p.Fin, fi = rem - h2 * p.Fin;
        return h == h2 ? __ldg(p.W + ((long long)fi * p.K + k) * p.Fout + fo) : 0.f;
    }
}

template <int MODE>
__global__ void __launch_bounds__(256) pack_w_bf16_kernel(const __grid_constant__ RowGemmBfParams p, uint8_t* __restrict__ image)
{
    const int per_blk = p.Ntot * p.kc;
    const int total = p.nkb * per_blk;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int j = idx / per_blk;
        const int rem = idx - j * per_blk;
        const int n = rem / p.kc, kk = rem - n * p.kc;
        const uint32_t off = (uint32_t)j * (uint32_t)(p.Ntot * p.rowbytes) + (uint32_t)n * p.rowbytes +
                             (uint32_t)(((kk >> 3) ^ swizzle_xor(n, p.swz)) * 16 + (kk & 7) * 2);
        *reinterpret_cast<bf16*>(image + off) = __float2bfloat16_rn(w_bf_value<MODE>(p, j, n, kk));
    }
}

// Column sums over the 32 rows held by the lanes of a warp (see contract_tc.cu): afterwards lane l holds column l in v[0].
__device__ __forceinline__ void warp_column_sums32(float (&v)[32], int lane)
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

template <int MODE, bool SUMS, bool POOL>
__global__ void __launch_bounds__(kBfThreads, 1)
rowgemm_bf16_kernel(const __grid_constant__ RowGemmBfParams p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t kb_bytes = (uint32_t)p.TM * kBM * p.rowbytes;            // one k-block of a stage
    const uint32_t stage_bytes = (uint32_t)p.nkb * kb_bytes;
    const uint32_t wblk_bytes = (uint32_t)p.Ntot * p.rowbytes;
    const uint32_t wbytes = ((uint32_t)p.nkb * wblk_bytes + 1023u) & ~1023u;
    const int S = p.nstages;

    uint8_t* sW = smem;
    uint8_t* sA = sW + wbytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)S * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kBfMaxStages;
    uint64_t* acc_full = bars + 2 * kBfMaxStages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* red = reinterpret_cast<float*>(tmem_slot + 2);                   // [2][256] batch-norm partial sums

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(acc_full + a, 1); mbar_init(acc_empty + a, 128); }
        fence_barrier_init();
        prefetch_tmap(&p.tmA0);
        if (MODE == BF_FWD && p.K > 1) prefetch_tmap(&p.tmAk);
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    if (SUMS) for (int i = threadIdx.x; i < 2 * 256; i += kBfThreads) red[i] = 0.f;
    {   // the weight image stays resident for the whole kernel
        const int nvec = (int)((uint32_t)p.nkb * wblk_bytes / 16);
        const uint4* src = reinterpret_cast<const uint4*>(p.wimage);
        uint4* dst = reinterpret_cast<uint4*>(sW);
        for (int i = threadIdx.x; i < nvec; i += kBfThreads) dst[i] = __ldg(src + i);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int acc_cols = p.TM * p.Ntot;                                     // TMEM columns of one accumulator stage

    if (warp == 0) {
        // ======================= TMA producer =======================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
                const int row0 = t * p.TM * kBM;
                mbar_wait(empty + s, ph ^ 1u, 21);
                mbar_expect_tx(full + s, stage_bytes);
                uint8_t* dst = sA + (size_t)s * stage_bytes;
                if (MODE == BF_FWD) {
                    for (int c = 0; c < p.nchunk; ++c) {
                        tma_load_2d(dst + (size_t)(c * p.K) * kb_bytes, &p.tmA0, full + s, c * p.kc, row0);
                        if (p.K > 1) tma_load_3d(dst + (size_t)(c * p.K + 1) * kb_bytes, &p.tmAk, full + s, c * p.kc, row0, 0);
                    }
                } else {
                    for (int j = 0; j < p.nkb; ++j) tma_load_2d(dst + (size_t)j * kb_bytes, &p.tmA0, full + s, j * p.kc, row0);
                }
                if (++s == S) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(kBM, p.Nsub, 0, 0);
            const uint32_t sbo = 8u * p.rowbytes;
            const int ksteps = p.kc / 16;
            const uint64_t dA0 = make_smem_desc(smem_u32(sA), 16, sbo, p.swz);
            const uint64_t dB0 = make_smem_desc(smem_u32(sW), 16, sbo, p.swz);
            const uint32_t a_stage = stage_bytes >> 4, a_kb = kb_bytes >> 4, a_m = ((uint32_t)kBM * p.rowbytes) >> 4;
            const uint32_t b_blk = wblk_bytes >> 4, b_sub = ((uint32_t)p.Nsub * p.rowbytes) >> 4;
            int s = 0, ti = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
                const int a = p.acc_stages == 2 ? (ti & 1) : 0;
                const uint32_t use = (uint32_t)(p.acc_stages == 2 ? (ti >> 1) : ti);
                mbar_wait(acc_empty + a, (use & 1) ^ 1u, 22);
                mbar_wait(full + s, ph, 23);
                tc_fence_after();
                for (int m = 0; m < p.TM; ++m) {
                    const uint32_t d0 = tmem_base + (uint32_t)(a * acc_cols + m * p.Ntot);
                    const uint64_t dAm = dA0 + (uint64_t)(a_stage * (uint32_t)s + a_m * (uint32_t)m);
                    for (int j = 0; j < p.nkb; ++j) {
                        const uint64_t dA = dAm + (uint64_t)(a_kb * (uint32_t)j);
                        const uint64_t dB = dB0 + (uint64_t)(b_blk * (uint32_t)j);
#pragma unroll 4
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint32_t acc = (j > 0 || ks > 0) ? 1u : 0u;
                            for (int sp = 0; sp < p.nsplit; ++sp)
                                mma_bf16(d0 + (uint32_t)(sp * p.Nsub), dA + 2u * ks, dB + (uint64_t)(b_sub * (uint32_t)sp) + 2u * ks, idesc, acc);
                        }
                    }
                }
                tc_commit(empty + s);                      // the stage is free once the MMAs have read it
                tc_commit(acc_full + a);
                if (++s == S) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ======================= epilogue: two groups of four warps take alternate tiles =======================
        const int grp = (warp - 2) >> 2;
        const int q = warp & 3;                                     // TMEM lane quarter this warp may read
        int ti = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
            const int a = p.acc_stages == 2 ? (ti & 1) : 0;
            if ((p.acc_stages == 2 ? a : 0) != grp) continue;
            const uint32_t use = (uint32_t)(p.acc_stages == 2 ? (ti >> 1) : ti);
            mbar_wait(acc_full + a, use & 1, 24);
            tc_fence_after();
            for (int m = 0; m < p.TM; ++m) {
                const uint32_t tbase = tmem_base + (uint32_t)(a * acc_cols + m * p.Ntot) + ((uint32_t)(q * 32) << 16);
                const long long prow = ((long long)t * p.TM + m) * kBM + q * 32 + lane;
                const bool ok = prow < p.Rp;
                if (MODE == BF_FWD) {
                    bf16* orow = p.out + prow * p.Nreal;
                    const bool wide = (p.Nreal & 15) == 0 && (((uintptr_t)p.out) & 31) == 0;
                    const bool sty = !POOL || p.out != nullptr;
                    for (int c0 = 0; c0 < p.Nreal; c0 += 32) {
                        float v[32];
                        tmem_ld16(tbase + c0, v);
                        if (c0 + 16 < p.Ntot) tmem_ld16(tbase + c0 + 16, v + 16);
                        tmem_ld_wait();
                        if (c0 + 16 >= p.Ntot) {
#pragma unroll
                            for (int i = 16; i < 32; ++i) v[i] = 0.f;
                        }
                        uint32_t w[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
                        if (ok && sty && wide) {
                            st_b32x8(orow + c0, w);
                            if (c0 + 16 < p.Nreal) st_b32x8(orow + c0 + 16, w + 8);
                        } else if (ok && sty) {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (c0 + i < p.Nreal) st1(orow + c0 + i, (i & 1) ? bf16_hi(w[i >> 1]) : bf16_lo(w[i >> 1]));
                        }
                        if (SUMS) {
                            // statistics of the STORED (rounded) tensor; rows past R and columns past Nreal are exact zeros
                            float sq[32];
#pragma unroll
                            for (int i = 0; i < 32; ++i) { v[i] = (i & 1) ? bf16_hi(w[i >> 1]) : bf16_lo(w[i >> 1]); sq[i] = v[i] * v[i]; }
                            warp_column_sums32(v, lane);
                            warp_column_sums32(sq, lane);
                            if (c0 + lane < p.Nreal) {
                                const int f = (c0 + lane) % p.Fout;
                                atomicAdd(red + f, v[0]);
                                atomicAdd(red + 256 + f, sq[0]);
                            }
                        }
                    }
                    if (POOL) {
                        // HEALPix max-pool over 4 consecutive rows, fused: batch norm without scale / shift, the bias and
                        // every activation of the library are non-decreasing per feature, so pooling the raw y first gives the same
                        // result (models.py:1235-1243) and the normalise / activate pass runs on a quarter of the rows.  A packed
                        // row holds rp consecutive rows: maximum over its rp column groups, then over the 4 / rp neighbouring lanes.
                        const int lpg = 4 / p.rp;
                        for (int f0 = 0; f0 < p.Fout; f0 += 16) {
                            float m[16];
                            tmem_ld16(tbase + f0, m);
                            tmem_ld_wait();
                            for (int h = 1; h < p.rp; ++h) {
                                float u[16];
                                tmem_ld16(tbase + h * p.Fout + f0, u);
                                tmem_ld_wait();
#pragma unroll
                                for (int i = 0; i < 16; ++i) m[i] = fmaxf(m[i], u[i]);
                            }
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                m[i] = fmaxf(m[i], __shfl_xor_sync(0xffffffffu, m[i], 1));
                                if (lpg == 4) m[i] = fmaxf(m[i], __shfl_xor_sync(0xffffffffu, m[i], 2));
                            }
                            if (ok && (lane & (lpg - 1)) == 0) {
                                uint32_t w[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(m[2 * i], m[2 * i + 1]);   // rounding is monotone: = max of the stored y
                                st_b32x8(p.pool_out + (prow / lpg) * p.Fout + f0, w);
                            }
                        }
                    }
                } else {
                    const int per = p.rp * p.Fin;                               // columns of one output plane of a packed row
                    const long long plane = p.Rp * (long long)per;
                    const bool wide = (((uintptr_t)p.out | (uintptr_t)(plane * 2)) & 31) == 0;
                    for (int c0 = 0; c0 < p.Nreal; c0 += 16) {
                        float v[16];
                        tmem_ld16(tbase + c0, v);
                        tmem_ld_wait();
                        if (ok) {
                            const int k = c0 / per, off = c0 - k * per;
                            bf16* o = p.out + k * plane + prow * per + off;
                            uint32_t w[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
                            if (wide) st_b32x8(o, w);
                            else { *reinterpret_cast<uint4*>(o) = make_uint4(w[0], w[1], w[2], w[3]);
                                   *reinterpret_cast<uint4*>(o + 8) = make_uint4(w[4], w[5], w[6], w[7]); }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(acc_empty + a);
        }
        if (SUMS) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            for (int e = threadIdx.x - 64; e < p.Fout; e += 256) {
                atomicAdd(p.bn_sums + e, (double)red[e]);
                atomicAdd(p.bn_sums + p.Fout + e, (double)red[256 + e]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

struct RowGemmBfPlan {
    int rp, rowbytes, kc, nkb, nchunk, Ntot, Nsub, nsplit, Nreal, swz, TM, nstages, acc_stages, tmem_cols;
    size_t smem, wbytes;
};

// red = reduction elements per row of the A operand (Fin forward, Fout data gradient); ncols_per_row = output columns produced
// per ORIGINAL row (Fout forward, K * Fin data gradient); planes = K forward, 1 data gradient
static bool plan_rowgemm_bf16(int mode, long long R, int red, int planes, int ncols_per_row, RowGemmBfPlan* pl)
{
    if (red >= 64) { if (red % 64) return false; pl->rp = 1; pl->rowbytes = 128; pl->nchunk = red / 64; }
    else if (red == 32) { pl->rp = 1; pl->rowbytes = 64; pl->nchunk = 1; }
    else if (red == 16) { pl->rp = 2; pl->rowbytes = 64; pl->nchunk = 1; }
    else return false;
    if (R % pl->rp) return false;
    pl->kc = pl->rowbytes / 2;
    pl->swz = pl->rowbytes == 128 ? 3 : 2;
    pl->nkb = planes * pl->nchunk;
    pl->Nreal = pl->rp * ncols_per_row;
    const int n16 = (pl->Nreal + 15) / 16 * 16;
    pl->nsplit = (n16 + 255) / 256;
    pl->Nsub = ((n16 + pl->nsplit - 1) / pl->nsplit + 15) / 16 * 16;
    pl->Ntot = pl->nsplit * pl->Nsub;
    if (pl->Ntot > 512) return false;
    pl->TM = (mode == BF_BWD_DATA && pl->rowbytes == 64 && 2 * pl->Ntot <= 512) ? 2 : 1;
    if (R / pl->rp < (long long)pl->TM * kBM) return false;
    pl->acc_stages = (2 * pl->TM * pl->Ntot <= 512) ? 2 : 1;
    pl->tmem_cols = 32;
    while (pl->tmem_cols < pl->acc_stages * pl->TM * pl->Ntot) pl->tmem_cols <<= 1;
    pl->wbytes = ((size_t)pl->nkb * pl->Ntot * pl->rowbytes + 1023) & ~(size_t)1023;
    const size_t stage = (size_t)pl->nkb * pl->TM * kBM * pl->rowbytes;
    const size_t fixed = 1024 + (2 * kBfMaxStages + 4) * 8 + 16 + 2 * 256 * 4 + 64;
    const size_t budget = 227 * 1024;
    if (pl->wbytes + fixed + 2 * stage > budget) return false;
    size_t ns = (budget - pl->wbytes - fixed) / stage;
    if (ns > kBfMaxStages) ns = kBfMaxStages;
    pl->nstages = (int)ns;
    pl->smem = pl->wbytes + fixed + ns * stage;
    return true;
}

template <int MODE, bool SUMS, bool POOL>
static int launch_rowgemm_bf16(RowGemmBfParams& p, size_t smem, uint8_t* image, cudaStream_t s)
{
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(rowgemm_bf16_kernel<MODE, SUMS, POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    p.wimage = image;
    const int total = p.nkb * p.Ntot * p.kc;
    pack_w_bf16_kernel<MODE><<<(total + 255) / 256, 256, 0, s>>>(p, image);
    count_launch();
    const int grid = p.ntiles < kNumSMs ? p.ntiles : kNumSMs;
    rowgemm_bf16_kernel<MODE, SUMS, POOL><<<grid, kBfThreads, smem, s>>>(p);
    DSB_LAUNCH_CHECK();
}

static void fill_params(RowGemmBfParams& p, const RowGemmBfPlan& pl, long long R, int Fin, int K, int Fout)
{
    p.Rp = R / pl.rp; p.Fin = Fin; p.K = K; p.Fout = Fout; p.rp = pl.rp;
    p.rowbytes = pl.rowbytes; p.kc = pl.kc; p.nkb = pl.nkb; p.nchunk = pl.nchunk;
    p.Ntot = pl.Ntot; p.Nsub = pl.Nsub; p.nsplit = pl.nsplit; p.Nreal = pl.Nreal; p.swz = pl.swz; p.TM = pl.TM;
    p.nstages = pl.nstages; p.acc_stages = pl.acc_stages; p.tmem_cols = pl.tmem_cols;
    p.ntiles = (int)((p.Rp + (long long)pl.TM * kBM - 1) / ((long long)pl.TM * kBM));
}

int contract_fwd_bf16(const bf16* x0, const bf16* xk, const float* W, bf16* Y, bf16* Ypool, long long R, int Fin, int K, int Fout,
                      double* bn_sums, void* workspace, cudaStream_t s, bool query, bool pool)
{
    if (R > 0x7fffff00LL || Fout > 256 || K < 1) return 1;
    RowGemmBfPlan pl;
    if (!plan_rowgemm_bf16(BF_FWD, R, Fin, K, Fout, &pl)) return 1;
    if (pool && ((Fout & 15) || (R & 3) || pl.rp > 2)) return 1;
    if (query) return 0;
    if ((((uintptr_t)Ypool) & 31) != 0 || (!Y && !Ypool)) return 1;
    if ((((uintptr_t)x0 | (uintptr_t)xk | (uintptr_t)Y | (uintptr_t)workspace) & 15) != 0 || !workspace) return 1;
    RowGemmBfParams p;
    fill_params(p, pl, R, Fin, K, Fout);
    const cuuint64_t cols = (cuuint64_t)pl.rp * Fin;
    {
        const cuuint64_t dims[2] = {cols, (cuuint64_t)p.Rp};
        const cuuint64_t strides[1] = {cols * 2};
        const cuuint32_t box[2] = {(cuuint32_t)pl.kc, (cuuint32_t)(pl.TM * kBM)};
        if (make_tmap_bf16(&p.tmA0, x0, 2, dims, strides, box, pl.swz)) return 1;
    }
    if (K > 1) {
        const cuuint64_t dims[3] = {cols, (cuuint64_t)p.Rp, (cuuint64_t)(K - 1)};
        const cuuint64_t strides[2] = {cols * 2, (cuuint64_t)R * Fin * 2};
        const cuuint32_t box[3] = {(cuuint32_t)pl.kc, (cuuint32_t)(pl.TM * kBM), (cuuint32_t)(K - 1)};
        if (make_tmap_bf16(&p.tmAk, xk, 3, dims, strides, box, pl.swz)) return 1;
    } else p.tmAk = p.tmA0;
    p.W = W; p.out = Y; p.pool_out = Ypool; p.bn_sums = bn_sums;
    if (Ypool) {
        if (!bn_sums) return launch_rowgemm_bf16<BF_FWD, false, true>(p, pl.smem, (uint8_t*)workspace, s);
        return launch_rowgemm_bf16<BF_FWD, true, true>(p, pl.smem, (uint8_t*)workspace, s);
    }
    if (!bn_sums) return launch_rowgemm_bf16<BF_FWD, false, false>(p, pl.smem, (uint8_t*)workspace, s);
    return launch_rowgemm_bf16<BF_FWD, true, false>(p, pl.smem, (uint8_t*)workspace, s);
}

int contract_bwd_data_bf16(const bf16* dY, const float* W, bf16* G, long long R, int Fin, int K, int Fout,
                           void* workspace, cudaStream_t s, bool query)
{
    if (R > 0x7fffff00LL || K < 1) return 1;
    RowGemmBfPlan pl;
    if (!plan_rowgemm_bf16(BF_BWD_DATA, R, Fout, 1, K * Fin, &pl)) return 1;
    if ((pl.rp * Fin) % 16) return 1;                                // 32-byte pieces of an output plane row
    if (query) return 0;
    if ((((uintptr_t)dY | (uintptr_t)G | (uintptr_t)workspace) & 15) != 0 || !workspace) return 1;
    RowGemmBfParams p;
    fill_params(p, pl, R, Fin, K, Fout);
    const cuuint64_t cols = (cuuint64_t)pl.rp * Fout;
    const cuuint64_t dims[2] = {cols, (cuuint64_t)p.Rp};
    const cuuint64_t strides[1] = {cols * 2};
    const cuuint32_t box[2] = {(cuuint32_t)pl.kc, (cuuint32_t)(pl.TM * kBM)};
    if (make_tmap_bf16(&p.tmA0, dY, 2, dims, strides, box, pl.swz)) return 1;
    p.tmAk = p.tmA0;
    p.W = W; p.out = G; p.pool_out = nullptr; p.bn_sums = nullptr;
    return launch_rowgemm_bf16<BF_BWD_DATA, false, false>(p, pl.smem, (uint8_t*)workspace, s);
}

size_t contract_workspace_bytes_bf16(long long R, int Fin, int K, int Fout)
{
    size_t need = 1024;
    RowGemmBfPlan pl;
    if (plan_rowgemm_bf16(BF_FWD, R, Fin, K, Fout, &pl) && pl.wbytes > need) need = pl.wbytes;
    if (plan_rowgemm_bf16(BF_BWD_DATA, R, Fout, 1, K * Fin, &pl) && pl.wbytes > need) need = pl.wbytes;
    return need + 1024;
}

// =====================================================================================================
// Weight gradient.  The reduction runs over the rows, so both operands are MN-major; for 16-bit elements that is the plain
// 128-byte swizzle: a TMA box of RS rows x 64 columns lands as the canonical layout (64 contiguous elements along M / N,
// 8-row groups along K 1024 bytes apart).  rp = 64 / Fin consecutive rows are viewed as one row of 64 features (and of
// rp*Fout gradient columns): no zero padding anywhere; of the rp x rp blocks of the product the rp diagonal ones add up to dW.
// Per stage: ONE box for plane 0, ONE rank-3 box for planes 1..K-1, ONE (rank-3) box for dY.  Every CTA accumulates a
// contiguous range of rows in TMEM and adds its partial result to dW with vector reductions at the end.
constexpr int kWgBfThreads = 192;                // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue

struct alignas(64) WGradBfParams {
    CUtensorMap tmX0, tmXk, tmdY;
    float* dW;
    long long Rp;
    int Fin, K, Fout, rp;
    int nblkA, nA, nB, mtiles;
    int RS, nstages, tmem_cols, vec_red, dy_rank3;
    long long iters_total, iters_per_cta;
};

// MN-major bf16 operand, 128-byte swizzle: LBO = distance between 64-element blocks along M / N, SBO = 8-row groups along K
__device__ __forceinline__ uint64_t make_mn_desc_bf16(uint32_t saddr, uint32_t block_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((block_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((1024u >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(kWgBfThreads, 1)
wgrad_bf16_kernel(const __grid_constant__ WGradBfParams p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.nstages;
    const uint32_t blk = (uint32_t)p.RS * 128u;                       // one 64-column block of RS rows
    const uint32_t stage_bytes = blk * (uint32_t)(p.nA + p.nB);       // [X blocks, chunk-major][dY blocks]
    const int N = p.nB * 64;

    uint8_t* sA = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)S * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kBfMaxStages;
    uint64_t* acc_full = bars + 2 * kBfMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
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

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < niter; ++i) {
                const int row0 = (int)((it0 + i) * p.RS);
                mbar_wait(empty + s, ph ^ 1u, 31);
                mbar_expect_tx(full + s, stage_bytes);
                uint8_t* dst = sA + (size_t)s * stage_bytes;
                for (int c = 0; c < p.nblkA; ++c) {
                    tma_load_2d(dst + (size_t)(c * p.K) * blk, &p.tmX0, full + s, c * 64, row0);
                    if (p.K > 1) tma_load_3d(dst + (size_t)(c * p.K + 1) * blk, &p.tmXk, full + s, c * 64, row0, 0);
                }
                uint8_t* dy = dst + (size_t)p.nA * blk;
                if (p.dy_rank3) tma_load_3d(dy, &p.tmdY, full + s, 0, row0, 0);
                else for (int c = 0; c < p.nB; ++c) tma_load_2d(dy + (size_t)c * blk, &p.tmdY, full + s, c * 64, row0);
                if (++s == S) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && niter > 0) {
            const uint32_t idesc = make_idesc_bf16(kBM, N, 1, 1);
            const uint64_t dA0 = make_mn_desc_bf16(smem_u32(sA), blk);
            const uint64_t dB0 = make_mn_desc_bf16(smem_u32(sA) + (uint32_t)p.nA * blk, blk);
            const uint32_t kstep = 2048u >> 4;                                  // 16 rows of 128 bytes
            const uint32_t mtA = (2u * blk) >> 4;                               // next 128 values of q
            const uint32_t stg = stage_bytes >> 4;
            const int ksteps = p.RS / 16;
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < niter; ++i) {
                mbar_wait(full + s, ph, 33);
                tc_fence_after();
                const uint64_t dA = dA0 + (uint64_t)(stg * (uint32_t)s), dB = dB0 + (uint64_t)(stg * (uint32_t)s);
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint32_t acc = (i > 0 || ks > 0) ? 1u : 0u;
                    for (int mt = 0; mt < p.mtiles; ++mt)
                        mma_bf16(tmem_base + (uint32_t)(mt * N), dA + (uint64_t)(kstep * (uint32_t)ks + mtA * (uint32_t)mt),
                                 dB + (uint64_t)(kstep * (uint32_t)ks), idesc, acc);
                }
                tc_commit(empty + s);
                if (++s == S) { s = 0; ph ^= 1u; }
            }
            tc_commit(acc_full);
        }
    } else if (niter > 0) {
        // epilogue, once per CTA: lane = value q of the padded [X_0 .. X_(K-1)] blocks, columns = packed gradient columns
        const int qd = warp & 3;
        mbar_wait(acc_full, 0, 35);
        tc_fence_after();
        for (int mt = 0; mt < p.mtiles; ++mt) {
            const int q = mt * kBM + qd * 32 + lane;
            const int bq = q >> 6, cq = q & 63;
            const int c = bq / p.K, k = bq - c * p.K;                // block index = chunk * K + plane
            const int pc = c * 64 + cq;                              // column of the packed plane row
            const int h = pc / p.Fin, fi = pc - h * p.Fin;           // sub-row, feature
            const bool ok = bq < p.nA;
            float* dst = p.dW + ((long long)fi * p.K + k) * p.Fout;
            const uint32_t tbase = tmem_base + (uint32_t)(mt * N) + ((uint32_t)(qd * 32) << 16);
            for (int c0 = 0; c0 < p.Fout; c0 += 8) {
                for (int hh = 0; hh < p.rp; ++hh) {                  // TMEM loads are warp-uniform: read every sub-row block
                    float v[8];
                    tmem_ld8(tbase + hh * p.Fout + c0, v);
                    tmem_ld_wait();
                    if (ok && hh == h) {
                        if (p.vec_red) { red_add_v4(dst + c0, v[0], v[1], v[2], v[3]); red_add_v4(dst + c0 + 4, v[4], v[5], v[6], v[7]); }
                        else {
#pragma unroll
                            for (int e = 0; e < 8; ++e) atomicAdd(dst + c0 + e, v[e]);
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

int contract_bwd_weight_bf16(const bf16* x0, const bf16* xk, const bf16* dY, float* dW, long long R, int Fin, int K, int Fout,
                             cudaStream_t s, bool query)
{
    if (R > 0x7fffff00LL || K < 1 || (Fout & 7)) return 1;
    WGradBfParams p;
    int rp = 1;
    if (Fin < 64) { if (64 % Fin) return 1; rp = 64 / Fin; }
    else if (Fin % 64) return 1;
    if (R % rp || (rp * Fout) % 64 || rp * Fout > 256) return 1;
    p.rp = rp; p.Fin = Fin; p.K = K; p.Fout = Fout;
    p.Rp = R / rp;
    p.nblkA = rp * Fin / 64;
    p.nA = K * p.nblkA;
    p.nB = rp * Fout / 64;
    p.mtiles = (p.nA + 1) / 2;
    const int N = p.nB * 64;
    if (p.mtiles * N > 512) return 1;
    int tcols = 32;
    while (tcols < p.mtiles * N) tcols <<= 1;
    p.tmem_cols = tcols;
    // rows per stage: the rank-3 box of the K - 1 planes should carry >= 16 KB; 2-4 stages must fit
    const size_t fixed = 1024 + (2 * kBfMaxStages + 2) * 8 + 64;
    const size_t budget = 227 * 1024;
    int RS = 64;
    while (RS > 16 && fixed + 3 * (size_t)RS * 128 * (p.nA + p.nB) > budget) RS >>= 1;
    if (fixed + 2 * (size_t)RS * 128 * (p.nA + p.nB) > budget) return 1;
    if (p.Rp < RS) return 1;
    p.RS = RS;
    const size_t stage = (size_t)RS * 128 * (p.nA + p.nB);
    size_t ns = (budget - fixed) / stage;
    if (ns > kBfMaxStages) ns = kBfMaxStages;
    p.nstages = (int)ns;
    const size_t smem = fixed + ns * stage;
    if (query) return 0;
    if ((((uintptr_t)x0 | (uintptr_t)xk | (uintptr_t)dY) & 15) != 0) return 1;
    const cuuint64_t colsX = (cuuint64_t)rp * Fin, colsY = (cuuint64_t)rp * Fout;
    {
        const cuuint64_t dims[2] = {colsX, (cuuint64_t)p.Rp};
        const cuuint64_t strides[1] = {colsX * 2};
        const cuuint32_t box[2] = {64, (cuuint32_t)RS};
        if (make_tmap_bf16(&p.tmX0, x0, 2, dims, strides, box, 3)) return 1;
    }
    if (K > 1) {
        const cuuint64_t dims[3] = {colsX, (cuuint64_t)p.Rp, (cuuint64_t)(K - 1)};
        const cuuint64_t strides[2] = {colsX * 2, (cuuint64_t)R * Fin * 2};
        const cuuint32_t box[3] = {64, (cuuint32_t)RS, (cuuint32_t)(K - 1)};
        if (make_tmap_bf16(&p.tmXk, xk, 3, dims, strides, box, 3)) return 1;
    } else p.tmXk = p.tmX0;
    p.dy_rank3 = 0;
    if (p.nB > 1) {
        // column blocks of the packed dY row as a third dimension: one box lands as [block][RS rows][128 bytes]
        // (if the driver refuses the decreasing stride: one rank-2 box per column block)
        const cuuint64_t dims[3] = {64, (cuuint64_t)p.Rp, (cuuint64_t)p.nB};
        const cuuint64_t strides[2] = {colsY * 2, 128};
        const cuuint32_t box[3] = {64, (cuuint32_t)RS, (cuuint32_t)p.nB};
        if (make_tmap_bf16(&p.tmdY, dY, 3, dims, strides, box, 3) == 0) p.dy_rank3 = 1;
    }
    if (!p.dy_rank3) {
        const cuuint64_t dims[2] = {colsY, (cuuint64_t)p.Rp};
        const cuuint64_t strides[1] = {colsY * 2};
        const cuuint32_t box[2] = {64, (cuuint32_t)RS};
        if (make_tmap_bf16(&p.tmdY, dY, 2, dims, strides, box, 3)) return 1;
    }
    p.dW = dW;
    p.vec_red = (((uintptr_t)dW & 15) == 0 && (Fout & 3) == 0) ? 1 : 0;
    p.iters_total = (p.Rp + RS - 1) / RS;
    int grid = p.iters_total < kNumSMs ? (int)p.iters_total : kNumSMs;
    p.iters_per_cta = (p.iters_total + grid - 1) / grid;
    grid = (int)((p.iters_total + p.iters_per_cta - 1) / p.iters_per_cta);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    wgrad_bf16_kernel<<<grid, kWgBfThreads, smem, s>>>(p);
    DSB_LAUNCH_CHECK();
}

}  // namespace tc
}  // namespace dsb200

using namespace dsb200;

// 1 when the bf16 tensor-core kernel of direction `dir` (0 forward, 1 data gradient, 2 weight gradient) takes this shape
extern "C" int dsb200_contract_bf16_supported(int dir, long long R, int Fin, int K, int Fout)
{
    if (dir == 0) return tc::contract_fwd_bf16(nullptr, nullptr, nullptr, nullptr, nullptr, R, Fin, K, Fout, nullptr, nullptr, 0, true, false) == 0;
    if (dir == 3) return tc::contract_fwd_bf16(nullptr, nullptr, nullptr, nullptr, nullptr, R, Fin, K, Fout, nullptr, nullptr, 0, true, true) == 0;
    if (dir == 1) return tc::contract_bwd_data_bf16(nullptr, nullptr, nullptr, R, Fin, K, Fout, nullptr, 0, true) == 0;
    return tc::contract_bwd_weight_bf16(nullptr, nullptr, nullptr, nullptr, R, Fin, K, Fout, 0, true) == 0;
}

extern "C" long long dsb200_contract_workspace_bytes_bf16(long long R, int Fin, int K, int Fout)
{
    return (long long)tc::contract_workspace_bytes_bf16(R, Fin, K, Fout);
}

extern "C" int dsb200_cheb_contract_fwd_bf16(const void* x0, const void* xk, const float* W, void* y, long long R, int Fin,
                                             int K, int Fout, double* bn_sums, void* workspace, void* stream)
{
    DSB_REQUIRE(x0 && W && y && R > 0 && (K == 1 || xk), "cheb_contract_fwd_bf16: bad arguments");
    const int rc = tc::contract_fwd_bf16((const bf16*)x0, (const bf16*)xk, W, (bf16*)y, nullptr, R, Fin, K, Fout, bn_sums, workspace,
                                         (cudaStream_t)stream, false, false);
    DSB_REQUIRE(rc != 1, "cheb_contract_fwd_bf16: unsupported shape / alignment (see dsb200_contract_bf16_supported)");
    return rc;
}

extern "C" int dsb200_cheb_contract_fwd_pool4_bf16(const void* x0, const void* xk, const float* W, void* y, void* ypool, long long R,
                                                   int Fin, int K, int Fout, double* bn_sums, void* workspace, void* stream)
{
    DSB_REQUIRE(x0 && W && ypool && R > 0 && (K == 1 || xk), "cheb_contract_fwd_pool4_bf16: bad arguments");
    const int rc = tc::contract_fwd_bf16((const bf16*)x0, (const bf16*)xk, W, (bf16*)y, (bf16*)ypool, R, Fin, K, Fout, bn_sums,
                                         workspace, (cudaStream_t)stream, false, true);
    DSB_REQUIRE(rc != 1, "cheb_contract_fwd_pool4_bf16: unsupported shape / alignment (dsb200_contract_bf16_supported(3, ...))");
    return rc;
}

extern "C" int dsb200_cheb_contract_bwd_data_bf16(const void* dy, const float* W, void* G, long long R, int Fin, int K, int Fout,
                                                  void* workspace, void* stream)
{
    DSB_REQUIRE(dy && W && G && R > 0, "cheb_contract_bwd_data_bf16: bad arguments");
    const int rc = tc::contract_bwd_data_bf16((const bf16*)dy, W, (bf16*)G, R, Fin, K, Fout, workspace, (cudaStream_t)stream, false);
    DSB_REQUIRE(rc != 1, "cheb_contract_bwd_data_bf16: unsupported shape / alignment (see dsb200_contract_bf16_supported)");
    return rc;
}

extern "C" int dsb200_cheb_contract_bwd_weight_bf16(const void* x0, const void* xk, const void* dy, float* dW, long long R,
                                                    int Fin, int K, int Fout, void* stream)
{
    DSB_REQUIRE(x0 && dy && dW && R > 0 && (K == 1 || xk), "cheb_contract_bwd_weight_bf16: bad arguments");
    const int rc = tc::contract_bwd_weight_bf16((const bf16*)x0, (const bf16*)xk, (const bf16*)dy, dW, R, Fin, K, Fout,
                                                (cudaStream_t)stream, false);
    DSB_REQUIRE(rc != 1, "cheb_contract_bwd_weight_bf16: unsupported shape / alignment (see dsb200_contract_bf16_supported)");
    return rc;
}
