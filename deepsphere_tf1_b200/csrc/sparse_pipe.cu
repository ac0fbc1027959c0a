// K1, bulk-asynchronous pipelined form:  out[b] = alpha * L x[b] + beta * y[b] + gamma * z[b]  on [B, M, F] tensors.
// Same operator and call sites as dsb200_spmm_ell (reference deepsphere/models.py:1056-1061 and its autodiff).
//
// Why: the gather kernels (sparse.cu) are bound by the SM's load/store data pipe, not by HBM: 9 scattered 16-byte reads and
// 18 record words per 16 output bytes are ~93 L1 wavefronts per 512 output bytes, each costing ~2 cycles when an instruction
// needs several (tools/exp_spmm_limits.py: 52 us at cosmo layer 2 even when every gather hits L1; a plain 3-stream kernel
// of the same size takes 33 us).  This kernel cuts the wavefronts per output by 3 and takes the compulsory traffic off the
// load/store pipe:
//   * a work item = (row tile of TR consecutive rows, sample); the x rows of the tile, and the y / z rows, are contiguous
//     in memory: one cp.async.bulk (TMA, 1-D) each into a shared-memory stage, completion on an mbarrier;
//   * the tile's halo (rows referenced from outside the tile; host-built list, graph.py: PipePlan) follows with 16-byte
//     cp.async -- about a third of a tile in HEALPix nested (Morton) order; items are taken grid-strided in memory order, so
//     the halo rows are L2 hits on the tiles the neighbouring CTAs are loading;
//   * the tile's records are one more bulk copy per item (L2 hits);
//   * S = 3 stages per CTA, filled by FOUR producer warps, one per stream (a bulk copy costs its issuing warp ~0.3 us
//     whatever its size); 2 CTAs per SM;
//   * consumer warps read ONLY shared memory.  Dense-group form (default): a lane group owns 4 consecutive rows (a 2 x 2
//     block of the sphere) and reads each of their <= 16 DISTINCT neighbour rows once (16 reads instead of 36), then applies
//     the group's dense 4 x 16 weight block; column slots alternate in bank-line parity between neighbouring groups and odd
//     groups store their rows pairwise swapped, so every shared-memory / global instruction of a warp touches both halves
//     of its 128-byte lines (no bank conflicts at 64-byte rows).  Record form (one lane group per row, W gathers, the
//     arithmetic order of the gather kernels: bit-identical results) is kept for ELL width 7 and as the reference.
// Measured (tools/exp_timing_method.py, back to back over rotating operand sets): cosmo layer 2 60 -> 42 us, layer 3 35 -> 23 us.
// F in {16, 32, 64} (NCH = 4, 8, 16 chunks per row), ELL width 7 / 9; other shapes keep the gather kernels.
#include "common.cuh"
#include "bf16.cuh"
#include <stdlib.h>

namespace dsb200 {

#ifdef DSB200_EXPERIMENTS
#define PIPE_ABLATE(a) ((a).ablate)
#else
#define PIPE_ABLATE(a) 0
#endif

struct PipeArgs {
    const int2* rec;        // [T][W][TR]  (local column, value bits); local column < TR: tile row, else TR + halo slot
    const int32_t* ext;     // [T][HC]     global row of every halo slot (unused slots: the tile's first row)
    const int32_t* nh;      // [T]         halo rows of the tile
    const float4 *x, *y, *z;
    float4* out;
    int T, TR, HC, B, M, S;
    int order;              // item order.  2 (default): grid-strided in memory order -- at any time the CTAs work on one window of
                            // consecutive row tiles of a sample, so halo rows are L2 hits on their neighbours' tiles (measured
                            // best); 0: CTA = contiguous item range, samples of a row tile consecutive; 1: contiguous range in
                            // memory order (experiments: DSB200_PIPE_ORDER)
    int ablate;             // timing experiments only (tools/exp_pipe_ablate.py): skip parts of the work
    int rev;                // walk the items from the END of memory to its start (same items, same results): a launch that starts
                            // where its producer stopped finds that producer's last ~100 MB in L2 (ops.py: zigzag recurrence)
    float alpha, beta, gamma;
};

__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(saddr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(saddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n .reg .pred p;\n W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D;\n bra W;\n D:\n}"
                 :: "r"(saddr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(saddr(dst)), "l"(src), "r"(bytes), "r"(saddr(bar)) : "memory");
}
__device__ __forceinline__ void cp16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(saddr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

constexpr int kConsumerWarps = 8;                                // record form: one lane group per row
constexpr int kProducerWarps = 4;                                // x + halo, records, y, z
constexpr int kDenseWarps = 4;                                   // dense-group form: one lane group per 4 rows
constexpr int kHaloPerLane = 16;                                 // halo chunks per producer lane: HC * NCH <= 512

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(saddr(bar)) : "memory");
}
// the mbarrier receives one arrival when all cp.async issued so far by this thread have landed (counted at init: .noinc)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(saddr(bar)) : "memory");
}

// Warp-specialised: warp 8 is the producer (it alone waits on global-memory latency: tile lists, bulk copies, halo
// copies), warps 0..7 consume -- each owns TR/8 rows of every item and synchronises with nobody but the producer
// (full[s] / empty[s] mbarriers), so the warps drift apart and hide each other's shared-memory latency.
// BF16: the tensors hold bfloat16 (csrc/bf16.cuh).  NCH stays the number of LANES per row (4 features each, now 8 bytes);
// the producers move CPR = NCH / 2 16-byte chunks per row; the arithmetic is fp32, the result is rounded once on the store.
template <bool BF16> struct PipeVec { typedef float4 type; };
template <> struct PipeVec<true> { typedef uint2 type; };
__device__ __forceinline__ float4 pv_load(const float4* p) { return *p; }
__device__ __forceinline__ float4 pv_load(const uint2* p) { return unpack_bf16x4(*p); }
__device__ __forceinline__ void pv_store(float4* p, float4 v) { *p = v; }
__device__ __forceinline__ void pv_store(uint2* p, float4 v) { *p = pack_bf16x4(v); }

template <int W, int NCH, bool HASY, bool HASZ, bool DENSE, bool BF16>
__global__ void __launch_bounds__((DENSE ? kDenseWarps : kConsumerWarps) * 32 + kProducerWarps * 32, 2)
spmm_pipe_kernel(const __grid_constant__ PipeArgs a)
{
    constexpr int CW = DENSE ? kDenseWarps : kConsumerWarps;
    constexpr int CPR = BF16 ? NCH / 2 : NCH;                      // 16-byte chunks per row
    typedef typename PipeVec<BF16>::type V;                        // what a lane reads / writes: its 4 features
    extern __shared__ __align__(128) unsigned char smem[];
    const int TR = a.TR, HC = a.HC, S = a.S;
    const uint32_t xbytes = (uint32_t)(TR + HC) * CPR * 16, tbytes = (uint32_t)TR * CPR * 16;
    const uint32_t rbytes = (uint32_t)W * TR * 8;
    const uint32_t ybase = xbytes, zbase = ybase + (HASY ? tbytes : 0), rbase = zbase + (HASZ ? tbytes : 0);
    const uint32_t sbytes = rbase + rbytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * sbytes);
    uint64_t* empty = full + S;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NBULK = 2 + (HASY ? 1 : 0) + (HASZ ? 1 : 0);

    const long long nitems = (long long)a.T * a.B;
    const long long c0 = nitems * blockIdx.x / gridDim.x, c1 = nitems * (blockIdx.x + 1) / gridDim.x;
    const long long i0 = a.order == 2 ? blockIdx.x : c0, istep = a.order == 2 ? gridDim.x : 1;
    const int n = a.order == 2 ? (int)((nitems - blockIdx.x + gridDim.x - 1) / gridDim.x) : (int)(c1 - c0);
    const bool rev = a.rev != 0;
    auto item_at = [&](int it) -> long long {
        const long long fwd = i0 + it * istep;
        return rev ? nitems - 1 - fwd : fwd;
    };
    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full + s, 32 + NBULK); mbar_init(empty + s, CW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= CW) {
        // ---------------- producers ----------------
        // One warp per stream (x + halo, records, y, z): a bulk copy costs its issuing warp ~0.3 us whatever its size
        // (profiles/r01_tma_box_rate.txt), so four copies per item from ONE warp would cap a CTA at one item per 1.3 us.
        const int role = warp - CW;                                           // 0: x + halo, 1: records, 2 / 3: y, z
        if (role == 2 && !HASY && !HASZ) return;
        if (role == 3 && !(HASY && HASZ)) return;
        if (role != 0 && lane != 0) return;
        // halo rows of the NEXT item are fetched (tile_ext, all HC slots: unused ones repeat the tile's first row) while
        // this one waits for its stage: no global-memory latency between a stage falling empty and its refill
        const int nhalo = (PIPE_ABLATE(a) & 2) ? 0 : HC * CPR;
        auto tile_of = [&](int it) -> int {
            const long long item = item_at(it);
            return a.order == 0 ? (int)(item / a.B) : (int)(item % a.T);
        };
        int gr[kHaloPerLane], grn[kHaloPerLane];
        if (role == 0) {
            const int32_t* ext = a.ext + (long long)tile_of(0) * HC;
#pragma unroll
            for (int q = 0; q < kHaloPerLane; ++q) { const int idx = lane + q * 32; grn[q] = idx < nhalo ? __ldg(ext + idx / CPR) : 0; }
        }
        for (int it = 0; it < n; ++it) {
            const int s = it % S, k = it / S;
            const long long item = item_at(it);
            int rt, b;
            if (a.order == 0) { rt = (int)(item / a.B); b = (int)(item - (long long)rt * a.B); }
            else { b = (int)(item / a.T); rt = (int)(item - (long long)b * a.T); }
            const int r0 = rt * TR, nt = min(TR, a.M - r0);
            unsigned char* st = smem + (size_t)s * sbytes;
            const long long g0 = ((long long)b * a.M + r0) * CPR;                 // first chunk of the tile in the sample
            const uint32_t nb = (uint32_t)nt * CPR * 16;
            if (role == 0) {
#pragma unroll
                for (int q = 0; q < kHaloPerLane; ++q) gr[q] = grn[q];
                if (it + 1 < n) {
                    const int32_t* ext = a.ext + (long long)tile_of(it + 1) * HC;
#pragma unroll
                    for (int q = 0; q < kHaloPerLane; ++q) { const int idx = lane + q * 32; grn[q] = idx < nhalo ? __ldg(ext + idx / CPR) : 0; }
                }
                if (k > 0) mbar_wait(empty + s, (uint32_t)(k - 1) & 1u);
                if (lane == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of this stage -> async writes
                    mbar_expect_tx(full + s, nb);
                    bulk_g2s(st, a.x + g0, nb, full + s);
                }
                float4* xh = reinterpret_cast<float4*>(st) + TR * CPR;
                const float4* xb = a.x + (long long)b * a.M * CPR;
#pragma unroll
                for (int q = 0; q < kHaloPerLane; ++q) {
                    const int idx = lane + q * 32;
                    if (idx < nhalo) cp16(xh + idx, xb + (long long)gr[q] * CPR + (idx % CPR));
                }
                cp_async_arrive(full + s);
            } else {
                if (k > 0) mbar_wait(empty + s, (uint32_t)(k - 1) & 1u);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (role == 1) {
                    mbar_expect_tx(full + s, rbytes);
                    bulk_g2s(st + rbase, a.rec + (long long)rt * W * TR, rbytes, full + s);
                } else if (role == 2) {
                    mbar_expect_tx(full + s, nb);
                    if (HASY) bulk_g2s(st + ybase, a.y + g0, nb, full + s); else bulk_g2s(st + zbase, a.z + g0, nb, full + s);
                } else {
                    mbar_expect_tx(full + s, nb);
                    bulk_g2s(st + zbase, a.z + g0, nb, full + s);
                }
            }
        }
        return;
    }

    if (DENSE) {
        // ---------------- consumers, dense-group form ----------------
        // A group = 4 consecutive rows (a 2 x 2 block of the sphere in nested order) and the U <= 16 DISTINCT columns its
        // 4 x W entries touch: a lane reads each distinct neighbour once (16 x 16 bytes instead of 36) and applies the
        // group's dense 4 x 16 weight block (zeros where a row does not see a column).  Records per tile: [18][G] 16-byte
        // pieces -- 16 weight pieces (row r ^ (g & 1), columns 4q..4q+3 at k = 4r + q) and 2 pieces of 8 uint16 local columns.
        // The pairwise row swap of odd groups and the parity-alternating column slots (graph.py) keep the 64-byte segments
        // of an instruction spread over both halves of the bank lines.
        constexpr int GPW = 32 / NCH;                                         // groups per warp pass
        const int G = TR / 4;
        const int sub = lane / NCH, lg = lane - sub * NCH;
        for (int it = 0; it < n; ++it) {
            const int s = it % S, k = it / S;
            const long long item = item_at(it);
            int rt, b;
            if (a.order == 0) { rt = (int)(item / a.B); b = (int)(item - (long long)rt * a.B); }
            else { b = (int)(item / a.T); rt = (int)(item - (long long)b * a.T); }
            const int r0 = rt * TR, nt = min(TR, a.M - r0);
            const unsigned char* st = smem + (size_t)s * sbytes;
            const V* xs = reinterpret_cast<const V*>(st);
            const V* ys = reinterpret_cast<const V*>(st + ybase);
            const V* zs = reinterpret_cast<const V*>(st + zbase);
            const float4* wrec = reinterpret_cast<const float4*>(st + rbase);   // [18][G]
            V* ob = reinterpret_cast<V*>(a.out) + ((long long)b * a.M + r0) * NCH;
            mbar_wait(full + s, (uint32_t)k & 1u);
            for (int g = warp * GPW + sub; g < G && 4 * g < nt; g += CW * GPW) {
                if (PIPE_ABLATE(a) & 16) {
                    for (int r = 0; r < 4; ++r) if (4 * g + r < nt && !(PIPE_ABLATE(a) & 8)) ob[(4 * g + r) * NCH + lg] = xs[(4 * g + r) * NCH + lg];
                    continue;
                }
                const uint4 c0 = *reinterpret_cast<const uint4*>(wrec + 16 * G + g);
                const uint4 c1 = *reinterpret_cast<const uint4*>(wrec + 17 * G + g);
                const uint32_t cw[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                float4 v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int c = (u & 1) ? (int)(cw[u >> 1] >> 16) : (int)(cw[u >> 1] & 0xffffu);
                    v[u] = pv_load(xs + c * NCH + lg);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 w = wrec[(4 * r + q) * G + g];
                        acc.x = fmaf(w.x, v[4 * q].x, acc.x); acc.y = fmaf(w.x, v[4 * q].y, acc.y); acc.z = fmaf(w.x, v[4 * q].z, acc.z); acc.w = fmaf(w.x, v[4 * q].w, acc.w);
                        acc.x = fmaf(w.y, v[4 * q + 1].x, acc.x); acc.y = fmaf(w.y, v[4 * q + 1].y, acc.y); acc.z = fmaf(w.y, v[4 * q + 1].z, acc.z); acc.w = fmaf(w.y, v[4 * q + 1].w, acc.w);
                        acc.x = fmaf(w.z, v[4 * q + 2].x, acc.x); acc.y = fmaf(w.z, v[4 * q + 2].y, acc.y); acc.z = fmaf(w.z, v[4 * q + 2].z, acc.z); acc.w = fmaf(w.z, v[4 * q + 2].w, acc.w);
                        acc.x = fmaf(w.w, v[4 * q + 3].x, acc.x); acc.y = fmaf(w.w, v[4 * q + 3].y, acc.y); acc.z = fmaf(w.w, v[4 * q + 3].z, acc.z); acc.w = fmaf(w.w, v[4 * q + 3].w, acc.w);
                    }
                    const int row = 4 * g + (r ^ (g & 1));                       // odd groups store their rows pairwise swapped
                    float4 res = make_float4(acc.x * a.alpha, acc.y * a.alpha, acc.z * a.alpha, acc.w * a.alpha);
                    if (HASY) { const float4 t = pv_load(ys + row * NCH + lg); res.x = fmaf(a.beta, t.x, res.x); res.y = fmaf(a.beta, t.y, res.y); res.z = fmaf(a.beta, t.z, res.z); res.w = fmaf(a.beta, t.w, res.w); }
                    if (HASZ) { const float4 t = pv_load(zs + row * NCH + lg); res.x = fmaf(a.gamma, t.x, res.x); res.y = fmaf(a.gamma, t.y, res.y); res.z = fmaf(a.gamma, t.z, res.z); res.w = fmaf(a.gamma, t.w, res.w); }
                    if (row < nt && !(PIPE_ABLATE(a) & 8)) pv_store(ob + row * NCH + lg, res);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
        return;
    }

    // ---------------- consumers ----------------
    constexpr int RPP = 32 / NCH;                                             // rows per warp pass
    const int rpw = TR / CW;                                      // rows per warp (host: TR % (8 * RPP) == 0)
    const int sub = lane / NCH, lg = lane - sub * NCH;
    for (int it = 0; it < n; ++it) {
        const int s = it % S, k = it / S;
        const long long item = item_at(it);
        int rt, b;
        if (a.order == 0) { rt = (int)(item / a.B); b = (int)(item - (long long)rt * a.B); }
        else { b = (int)(item / a.T); rt = (int)(item - (long long)b * a.T); }
        const int r0 = rt * TR, nt = min(TR, a.M - r0);
        const unsigned char* st = smem + (size_t)s * sbytes;
        const V* xs = reinterpret_cast<const V*>(st);
        const V* ys = reinterpret_cast<const V*>(st + ybase);
        const V* zs = reinterpret_cast<const V*>(st + zbase);
        const int2* srec = reinterpret_cast<const int2*>(st + rbase);
        V* ob = reinterpret_cast<V*>(a.out) + ((long long)b * a.M + r0) * NCH;
        mbar_wait(full + s, (uint32_t)k & 1u);
        const int rend = min(nt, (warp + 1) * rpw);
        for (int r = warp * rpw + sub; r < rend; r += RPP) {
            if (PIPE_ABLATE(a) & 16) { if (!(PIPE_ABLATE(a) & 8)) ob[r * NCH + lg] = xs[r * NCH + lg]; continue; }
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < W; ++j) {
                if ((PIPE_ABLATE(a) & 1) && j > 0) break;
                const int2 rc = srec[j * TR + r];
                const float w = __int_as_float(rc.y);
                const float4 v = pv_load(xs + rc.x * NCH + lg);
                acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
            }
            float4 res = make_float4(acc.x * a.alpha, acc.y * a.alpha, acc.z * a.alpha, acc.w * a.alpha);
            if (HASY) { const float4 t = pv_load(ys + r * NCH + lg); res.x = fmaf(a.beta, t.x, res.x); res.y = fmaf(a.beta, t.y, res.y); res.z = fmaf(a.beta, t.z, res.z); res.w = fmaf(a.beta, t.w, res.w); }
            if (HASZ) { const float4 t = pv_load(zs + r * NCH + lg); res.x = fmaf(a.gamma, t.x, res.x); res.y = fmaf(a.gamma, t.y, res.y); res.z = fmaf(a.gamma, t.z, res.z); res.w = fmaf(a.gamma, t.w, res.w); }
            if (!(PIPE_ABLATE(a) & 8)) pv_store(ob + r * NCH + lg, res);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
    }
}

static size_t pipe_smem(int W, int TR, int HC, int nch, int extra, int S) {
    const size_t stage = (size_t)(TR + HC) * nch * 16 + (size_t)TR * nch * 16 * extra + (size_t)W * TR * 8;
    return (size_t)S * stage + 2 * S * 8 + 64;
}

template <int W, int NCH, bool DENSE, bool BF16>
static int pipe_launch_t(const PipeArgs& a, int grid, size_t smem, cudaStream_t s)
{
    constexpr int threads = ((DENSE ? kDenseWarps : kConsumerWarps) + kProducerWarps) * 32;
#define DSB_PIPE(YY, ZZ) do { \
        static bool cfg = false; \
        if (!cfg) { cudaError_t e = cudaFuncSetAttribute(spmm_pipe_kernel<W, NCH, YY, ZZ, DENSE, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024); \
                    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return (int)e; } cfg = true; } \
        spmm_pipe_kernel<W, NCH, YY, ZZ, DENSE, BF16><<<grid, threads, smem, s>>>(a); } while (0)
    if (a.y && a.z) DSB_PIPE(true, true);
    else if (a.y) DSB_PIPE(true, false);
    else if (a.z) DSB_PIPE(false, true);
    else DSB_PIPE(false, false);
#undef DSB_PIPE
    DSB_LAUNCH_CHECK();
}

}  // namespace dsb200

using namespace dsb200;

template <bool BF16>
static int spmm_pipe_entry(const void* tile_recs, const int32_t* tile_ext, const int32_t* tile_nh,
                           int ntiles, int TR, int HC, int width, int dense, int B, int M, int F,
                           const void* x, const void* y, const void* z, void* out,
                           float alpha, float beta, float gamma, void* stream)
{
    DSB_REQUIRE(B > 0 && M > 0 && F > 0, "spmm_pipe: empty matrix");
    DSB_REQUIRE(tile_recs && tile_ext && tile_nh && x && out, "spmm_pipe: null pointer");
    DSB_REQUIRE(x != out, "spmm_pipe: out must not alias x");
    DSB_REQUIRE(width == 9 || width == 7, "spmm_pipe: ELL width must be 7 or 9");
    const int reverse = (dense >> 1) & 1;                       // bit 1 of `dense`: walk the items from the end of memory
    dense &= 1;
    DSB_REQUIRE(!dense || width == 9, "spmm_pipe: dense-group records are 72 bytes per row (width 9 layout)");
    DSB_REQUIRE(F == 16 || F == 32 || F == 64, "spmm_pipe: F must be 16, 32 or 64");
    DSB_REQUIRE(TR > 0 && HC >= 0 && ntiles == (M + TR - 1) / TR, "spmm_pipe: tile plan does not match M");
    const int nch = F / 4, cpr = BF16 ? nch / 2 : nch;          // lanes per row, 16-byte chunks per row
    DSB_REQUIRE(TR % (8 * (128 / F)) == 0 && HC * cpr <= 32 * kHaloPerLane && (width * TR * 8) % 16 == 0, "spmm_pipe: unsupported tile shape");
    if (beta == 0.f) y = nullptr;
    if (gamma == 0.f) z = nullptr;
    const uintptr_t align = (uintptr_t)x | (uintptr_t)out | (uintptr_t)y | (uintptr_t)z;
    DSB_REQUIRE((align & 15) == 0, "spmm_pipe: pointers must be 16-byte aligned");
    const int extra = (y ? 1 : 0) + (z ? 1 : 0);
    int S = 4;
    { const int e = DSB_ENV_INT("DSB200_PIPE_STAGES", 0); if (e >= 2 && e <= 4) S = e; }
    while (S > 2 && pipe_smem(width, TR, HC, cpr, extra, S) > 113 * 1024) --S;
    const size_t smem = pipe_smem(width, TR, HC, cpr, extra, S);
    DSB_REQUIRE(smem <= 113 * 1024, "spmm_pipe: tile does not fit in shared memory");
    PipeArgs a;
    a.rec = (const int2*)tile_recs; a.ext = tile_ext; a.nh = tile_nh;
    a.x = (const float4*)x; a.y = (const float4*)y; a.z = (const float4*)z; a.out = (float4*)out;
    a.T = ntiles; a.TR = TR; a.HC = HC; a.B = B; a.M = M; a.S = S;
    a.alpha = alpha; a.beta = beta; a.gamma = gamma;
    a.ablate = DSB_EXPERIMENT_INT("DSB200_PIPE_ABLATE", 0);
    a.order = DSB_ENV_INT("DSB200_PIPE_ORDER", 2);
    a.rev = reverse;
    int ctas = 2;
    if (DSB_ENV_INT("DSB200_PIPE_CTAS", 2) == 1) ctas = 1;
    const long long nitems = (long long)ntiles * B;
    const long long cap = (long long)ctas * kNumSMs;
    const int grid = (int)(nitems < cap ? nitems : cap);
    cudaStream_t s = (cudaStream_t)stream;
    if (dense) {                                                  // the records do not depend on the ELL width (W = 9 is a dummy)
        if (nch == 4) return pipe_launch_t<9, 4, true, BF16>(a, grid, smem, s);
        if (nch == 8) return pipe_launch_t<9, 8, true, BF16>(a, grid, smem, s);
        return pipe_launch_t<9, 16, true, BF16>(a, grid, smem, s);
    }
    if (width == 9) {
        if (nch == 4) return pipe_launch_t<9, 4, false, BF16>(a, grid, smem, s);
        if (nch == 8) return pipe_launch_t<9, 8, false, BF16>(a, grid, smem, s);
        return pipe_launch_t<9, 16, false, BF16>(a, grid, smem, s);
    }
    if (nch == 4) return pipe_launch_t<7, 4, false, BF16>(a, grid, smem, s);
    if (nch == 8) return pipe_launch_t<7, 8, false, BF16>(a, grid, smem, s);
    return pipe_launch_t<7, 16, false, BF16>(a, grid, smem, s);
}

extern "C" int dsb200_spmm_pipe(const void* tile_recs, const int32_t* tile_ext, const int32_t* tile_nh,
                                int ntiles, int TR, int HC, int width, int dense, int B, int M, int F,
                                const float* x, const float* y, const float* z, float* out,
                                float alpha, float beta, float gamma, void* stream)
{
    return spmm_pipe_entry<false>(tile_recs, tile_ext, tile_nh, ntiles, TR, HC, width, dense, B, M, F, x, y, z, out, alpha, beta, gamma, stream);
}

// the same operator on bfloat16 tensors (BF16 mode); the tile plan is the one of rows of 2 F bytes (graph.pipe_plan(.., esize=2))
extern "C" int dsb200_spmm_pipe_bf16(const void* tile_recs, const int32_t* tile_ext, const int32_t* tile_nh,
                                     int ntiles, int TR, int HC, int width, int dense, int B, int M, int F,
                                     const void* x, const void* y, const void* z, void* out,
                                     float alpha, float beta, float gamma, void* stream)
{
    return spmm_pipe_entry<true>(tile_recs, tile_ext, tile_nh, ntiles, TR, HC, width, dense, B, M, F, x, y, z, out, alpha, beta, gamma, stream);
}
