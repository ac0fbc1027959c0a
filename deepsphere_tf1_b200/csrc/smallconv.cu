// ChebConv layers with a tiny reduction depth Q = K*Fin <= 8 (the first layer of the cosmo and
// similar nets: Fin = 1, K = 5, Fout = 16 on 262,144 vertices x 16 samples).  There the filter output
// y [B, M, Fout] is Fout/Q times larger than the Chebyshev basis it is computed from, so it is never
// written to memory: every consumer recomputes  y[r, :] = sum_q X_q[r] W[q, :]  (Q FMAs per output)
// from the basis.
//   * batch-norm statistics (tf.nn.moments, models.py:1197-1203) come from the Q x Q Gram matrix of the
//     basis:  sum_r y_f = sum_q W_qf s_q,   sum_r y_f^2 = sum_qq' W_qf W_q'f G_qq'
//   * forward:  basis -> y -> normalise -> bias -> activation -> pool in one pass (models.py:1229-1243)
//   * backward: basis + dout -> gz (recomputed), S1 = sum gz, S2 = sum gz*xhat, A = X^T gz in one pass;
//     the batch-norm backward and the weight gradient X^T dy then follow in closed form from A, S1, S2
//     and the Gram matrix (no dy, no gz, no y in memory).
// Index conventions: q = k*Fin + fi inside this file; the weight row in W is fi*K + k (models.py:1062-1066).
#include "common.cuh"
#include <float.h>

namespace dsb200 {

constexpr int QMAX = 8;

struct Planes {
    const float* p0;
    const float* pk;
    long long plane_stride;      // R * Fin
    int Fin, K;
    // pointer to element (row 0, column q) of the concatenated basis; hoisted out of the row loops
    __device__ __forceinline__ const float* col(int q) const {
        const int k = q / Fin, fi = q - k * Fin;
        return (k == 0 ? p0 : pk + (long long)(k - 1) * plane_stride) + fi;
    }
};

// Row addressing of the basis planes.  Classic: row (b, m) = b*M + m.  Sample-interleaved (il != 0): the input
// was transposed to [M, B, Fin] before the recurrence (so that the sparse products gather B*Fin contiguous
// floats per neighbour instead of Fin), row (b, m) = m*B + b.  Work items t enumerate pooled rows with the
// sample index fastest in the interleaved case (coalesced basis reads); outputs keep the classic layout.
// one 32-byte read-only load (sm_100+), src 32-byte aligned
__device__ __forceinline__ void ldg8(const float* src, float* v) {
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(src));
}

struct RowMap {
    int B, M, Mo, p, il;
    // B * M < 2^31 (checked by the launchers): 32-bit index arithmetic
    __device__ __forceinline__ void decode(long long t64, int& b, int& mo) const {
        const unsigned t = (unsigned)t64;
        if (il) { mo = (int)(t / (unsigned)B); b = (int)(t - (unsigned)mo * (unsigned)B); }
        else { b = (int)(t / (unsigned)Mo); mo = (int)(t - (unsigned)b * (unsigned)Mo); }
    }
    __device__ __forceinline__ int xrow(int b, int mo, int j) const {
        const int m = mo * p + j;
        return il ? m * B + b : b * M + m;
    }
};

// gram[0..Q) = sum_r X_q[r];  gram[Q + q*Q + q'] = sum_r X_q[r] X_q'[r]   (double, caller zeroes)
template <int Q>
__global__ void __launch_bounds__(256)
gram_kernel(Planes P, long long R, double* __restrict__ gram)
{
    __shared__ float red[Q + Q * Q];
    for (int i = threadIdx.x; i < Q + Q * Q; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    float s[Q], g[Q * (Q + 1) / 2];
#pragma unroll
    for (int i = 0; i < Q; ++i) s[i] = 0.f;
#pragma unroll
    for (int i = 0; i < Q * (Q + 1) / 2; ++i) g[i] = 0.f;
    const float* cp[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) cp[q] = P.col(q);
    const int Fin = P.Fin;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < R; r += (long long)gridDim.x * blockDim.x) {
        float x[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) x[q] = __ldg(cp[q] + r * Fin);
        int t = 0;
#pragma unroll
        for (int a = 0; a < Q; ++a) {
            s[a] += x[a];
#pragma unroll
            for (int b = a; b < Q; ++b) { g[t] = fmaf(x[a], x[b], g[t]); ++t; }
        }
    }
    // warp reduce, then shared, then one double atomic per entry and block
    int t = 0;
#pragma unroll
    for (int a = 0; a < Q; ++a) {
        float v = s[a];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&red[a], v);
#pragma unroll
        for (int b = a; b < Q; ++b) {
            float w = g[t++];
            for (int o = 16; o; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
            if ((threadIdx.x & 31) == 0) { atomicAdd(&red[Q + a * Q + b], w); if (b != a) atomicAdd(&red[Q + b * Q + a], w); }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Q + Q * Q; i += blockDim.x) atomicAdd(gram + i, (double)red[i]);
}

// sums[f] = sum_q W_qf s_q ; sums[F + f] = sum_qq' W_qf W_q'f G_qq'
__global__ void bn_sums_from_gram_kernel(const double* __restrict__ gram, const float* __restrict__ W,
                                         int Fin, int K, int Fout, double* __restrict__ sums)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= Fout) return;
    const int Q = Fin * K;
    double w[QMAX];
    for (int q = 0; q < Q; ++q) { const int k = q / Fin, fi = q - k * Fin; w[q] = (double)W[(fi * K + k) * Fout + f]; }
    double s1 = 0.0, s2 = 0.0;
    for (int a = 0; a < Q; ++a) {
        s1 += w[a] * gram[a];
        for (int b = 0; b < Q; ++b) s2 += w[a] * w[b] * gram[Q + a * Q + b];
    }
    sums[f] = s1;
    sums[Fout + f] = s2;
}

// One thread = one pooled row x 4 output features.
template <int Q>
__global__ void __launch_bounds__(256)
smallconv_fwd_kernel(Planes P, RowMap rm, const float* __restrict__ W, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ bias, float* __restrict__ out,
                     long long Rout, int Fout, int p, int pool, int act)
{
    extern __shared__ float sm[];                    // Ws[Q][Fout], mu[Fout], rs[Fout], bs[Fout]
    float* Ws = sm;
    float* mu = sm + Q * Fout;
    float* rs = mu + Fout;
    float* bs = rs + Fout;
    for (int i = threadIdx.x; i < Q * Fout; i += blockDim.x) {
        const int q = i / Fout, f = i - q * Fout;
        const int k = q / P.Fin, fi = q - k * P.Fin;
        Ws[i] = __ldg(W + (fi * P.K + k) * Fout + f);
    }
    for (int f = threadIdx.x; f < Fout; f += blockDim.x) {
        mu[f] = mean ? __ldg(mean + f) : 0.f;
        rs[f] = rstd ? __ldg(rstd + f) : 1.f;
        bs[f] = bias ? __ldg(bias + f) : 0.f;
    }
    __syncthreads();
    const int Fv = Fout / 4;
    const long long total = Rout * Fv;
    const float* cp[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) cp[q] = P.col(q);
    const int Fin = P.Fin;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const long long t = g / Fv;
        const int f0 = (int)(g - t * Fv) * 4;
        int b, mo;
        rm.decode(t, b, mo);
        const long long ro = (long long)b * rm.Mo + mo;
        float acc[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[v] = (pool == DSB200_POOL_MAX) ? -FLT_MAX : 0.f;
        for (int j = 0; j < p; ++j) {
            const long long r = rm.xrow(b, mo, j);
            float y[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const float x = __ldg(cp[q] + r * Fin);
                const float4 w = *reinterpret_cast<const float4*>(Ws + q * Fout + f0);
                y[0] = fmaf(x, w.x, y[0]); y[1] = fmaf(x, w.y, y[1]); y[2] = fmaf(x, w.z, y[2]); y[3] = fmaf(x, w.w, y[3]);
            }
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const float a = act_fwd((y[v] - mu[f0 + v]) * rs[f0 + v] + bs[f0 + v], act);
                acc[v] = (pool == DSB200_POOL_MAX) ? fmaxf(acc[v], a) : acc[v] + a;
            }
        }
        if (pool != DSB200_POOL_MAX) {
            const float inv = 1.f / (float)p;
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[v] *= inv;
        }
        st4(out + ro * Fout + f0, make_float4(acc[0], acc[1], acc[2], acc[3]));
    }
}

// acc (double, zeroed): [0, Q*Fout) A[q][f] = sum_r X_q[r] gz[r,f];  then S1[f] = sum gz, S2[f] = sum gz*xhat
template <int Q>
__global__ void __launch_bounds__(256)
smallconv_bwd_kernel(Planes P, RowMap rm, const float* __restrict__ W, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ bias, const float* __restrict__ dout,
                     double* __restrict__ accg, long long Rout, int Fout, int p, int pool, int act)
{
    extern __shared__ float sm[];                    // Ws[Q][Fout], mu, rs, bs, red[(Q+2)*Fout]
    float* Ws = sm;
    float* mu = sm + Q * Fout;
    float* rs = mu + Fout;
    float* bs = rs + Fout;
    float* red = bs + Fout;
    for (int i = threadIdx.x; i < Q * Fout; i += blockDim.x) {
        const int q = i / Fout, f = i - q * Fout;
        const int k = q / P.Fin, fi = q - k * P.Fin;
        Ws[i] = __ldg(W + (fi * P.K + k) * Fout + f);
    }
    for (int f = threadIdx.x; f < Fout; f += blockDim.x) {
        mu[f] = mean ? __ldg(mean + f) : 0.f;
        rs[f] = rstd ? __ldg(rstd + f) : 1.f;
        bs[f] = bias ? __ldg(bias + f) : 0.f;
    }
    for (int i = threadIdx.x; i < (Q + 2) * Fout; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    const int Fv = Fout / 4;
    // threads of a block keep their feature group over the grid-stride loop: blockDim.x % Fv == 0
    const int rows_per_block = blockDim.x / Fv;
    const int f0 = (threadIdx.x % Fv) * 4, rl = threadIdx.x / Fv;
    float A[Q][4], s1[4], s2[4];
    const float* cp[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) { A[q][0] = A[q][1] = A[q][2] = A[q][3] = 0.f; cp[q] = P.col(q); }
    const int Fin = P.Fin;
#pragma unroll
    for (int v = 0; v < 4; ++v) { s1[v] = 0.f; s2[v] = 0.f; }
    for (long long t = (long long)blockIdx.x * rows_per_block + rl; t < Rout; t += (long long)gridDim.x * rows_per_block) {
        int b, mo;
        rm.decode(t, b, mo);
        const long long ro = (long long)b * rm.Mo + mo;
        const float4 d4 = ldg4(dout + ro * Fout + f0);
        const float d[4] = {d4.x, d4.y, d4.z, d4.w};
        float best[4], gsel[4], xhsel[4];
        int arg[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) { best[v] = -FLT_MAX; arg[v] = 0; gsel[v] = 0.f; xhsel[v] = 0.f; }
        const float inv = 1.f / (float)p;
        for (int j = 0; j < p; ++j) {
            const long long r = rm.xrow(b, mo, j);
            float x[Q], y[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                x[q] = __ldg(cp[q] + r * Fin);
                const float4 w = *reinterpret_cast<const float4*>(Ws + q * Fout + f0);
                y[0] = fmaf(x[q], w.x, y[0]); y[1] = fmaf(x[q], w.y, y[1]); y[2] = fmaf(x[q], w.z, y[2]); y[3] = fmaf(x[q], w.w, y[3]);
            }
            if (pool == DSB200_POOL_MAX && p > 1) {
                // remember the first arg-max sibling; its contribution is added after the loop
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float xh = (y[v] - mu[f0 + v]) * rs[f0 + v];
                    const float a = act_fwd(xh + bs[f0 + v], act);
                    if (a > best[v]) { best[v] = a; arg[v] = j; xhsel[v] = xh; gsel[v] = d[v] * act_grad(xh + bs[f0 + v], act); }
                }
            } else {
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float xh = (y[v] - mu[f0 + v]) * rs[f0 + v];
                    const float g = d[v] * inv * act_grad(xh + bs[f0 + v], act);
                    s1[v] += g; s2[v] = fmaf(g, xh, s2[v]);
#pragma unroll
                    for (int q = 0; q < Q; ++q) A[q][v] = fmaf(x[q], g, A[q][v]);
                }
            }
        }
        if (pool == DSB200_POOL_MAX && p > 1) {
#pragma unroll
            for (int v = 0; v < 4; ++v) { s1[v] += gsel[v]; s2[v] = fmaf(gsel[v], xhsel[v], s2[v]); }
            // X of the arg-max sibling of every feature (re-read: L1 hit)
#pragma unroll
            for (int q = 0; q < Q; ++q) {
#pragma unroll
                for (int v = 0; v < 4; ++v) A[q][v] = fmaf(__ldg(cp[q] + rm.xrow(b, mo, arg[v]) * Fin), gsel[v], A[q][v]);
            }
        }
    }
    const bool pw = Fv < 32 && (Fv & (Fv - 1)) == 0;
    const bool leader = !pw || (threadIdx.x & 31) < Fv;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const float a = pw ? warp_sum_mod(A[q][v], Fv) : A[q][v];
            if (leader) atomicAdd(&red[q * Fout + f0 + v], a);
        }
        const float a1 = pw ? warp_sum_mod(s1[v], Fv) : s1[v];
        const float a2 = pw ? warp_sum_mod(s2[v], Fv) : s2[v];
        if (leader) { atomicAdd(&red[Q * Fout + f0 + v], a1); atomicAdd(&red[(Q + 1) * Fout + f0 + v], a2); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (Q + 2) * Fout; i += blockDim.x) atomicAdd(accg + i, (double)red[i]);
}

// ---- fast path: ReLU + max pooling over 4 siblings (the configuration of every BASELINE classifier).
// One thread = one pooled row x 8 output features.  relu and max commute with the positive affine map of
// batch norm, so the forward keeps a running max of z = y*rstd + c and clamps once; it also records which
// sibling won (1 byte per output).  The backward then needs no recomputation of y at all: with
// out = relu(xhat + b) the gradient is gz = dout * (out > 0), xhat = out - b where it matters, and the
// arg-max byte selects the basis row for A = X^T gz.
template <int Q>
__global__ void __launch_bounds__(256)
smallconv_fwd_fast_kernel(Planes P, RowMap rm, const float* __restrict__ W, const float* __restrict__ mean,
                          const float* __restrict__ rstd, const float* __restrict__ bias, float* __restrict__ out,
                          uint8_t* __restrict__ argmax, long long Rout, int Fout)
{
    extern __shared__ float sm[];                    // Ws[Q][Fout], rs[Fout], cs[Fout]
    float* Ws = sm;
    float* rs = sm + Q * Fout;
    float* cs = rs + Fout;
    for (int i = threadIdx.x; i < Q * Fout; i += blockDim.x) {
        const int q = i / Fout, f = i - q * Fout;
        const int k = q / P.Fin, fi = q - k * P.Fin;
        Ws[i] = __ldg(W + (fi * P.K + k) * Fout + f);
    }
    for (int f = threadIdx.x; f < Fout; f += blockDim.x) {
        const float r = rstd ? __ldg(rstd + f) : 1.f;
        rs[f] = r;
        cs[f] = (bias ? __ldg(bias + f) : 0.f) - (mean ? __ldg(mean + f) : 0.f) * r;      // z = y*r + c
    }
    __syncthreads();
    const int Fg = Fout / 8;
    const long long total = Rout * Fg;
    const bool wide = ((uintptr_t)out & 31) == 0;        // Fout % 8 == 0 here: every lane's piece is 32-byte aligned
    const float* cp[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) cp[q] = P.col(q);
    const int Fin = P.Fin;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const unsigned gu = (unsigned)g;                 // Rout * Fg < 2^31
        const unsigned t = gu / (unsigned)Fg;
        const int f0 = (int)(gu - t * (unsigned)Fg) * 8;
        int b, mo;
        rm.decode(t, b, mo);
        float x[4][Q];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long r = (long long)rm.xrow(b, mo, j) * Fin;
#pragma unroll
            for (int q = 0; q < Q; ++q) x[j][q] = __ldg(cp[q] + r);
        }
        float y[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int v = 0; v < 8; ++v) y[j][v] = 0.f;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const float4 w0 = *reinterpret_cast<const float4*>(Ws + q * Fout + f0);
            const float4 w1 = *reinterpret_cast<const float4*>(Ws + q * Fout + f0 + 4);
            const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int v = 0; v < 8; ++v) y[j][v] = fmaf(x[j][q], w[v], y[j][v]);
        }
        float o[8];
        uint32_t a0 = 0, a1 = 0;
        // (vector reads of the scale / shift and one 32-byte store per lane: the kernel is bound by load/store-pipe
        // wavefronts -- a lane's output row piece lies in a different sample plane than its neighbours')
        const float4 r0 = *reinterpret_cast<const float4*>(rs + f0), r1 = *reinterpret_cast<const float4*>(rs + f0 + 4);
        const float4 c0 = *reinterpret_cast<const float4*>(cs + f0), c1 = *reinterpret_cast<const float4*>(cs + f0 + 4);
        const float rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
        const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            const float r = rr[v], c = cc[v];
            float best = fmaf(y[0][v], r, c);
            uint32_t arg = 0;
#pragma unroll
            for (int j = 1; j < 4; ++j) {
                const float z = fmaf(y[j][v], r, c);
                if (z > best) { best = z; arg = j; }
            }
            o[v] = fmaxf(best, 0.f);
            if (v < 4) a0 |= arg << (8 * v); else a1 |= arg << (8 * (v - 4));
        }
        const long long ro = (long long)b * rm.Mo + mo;
        if (wide) {
            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                         :: "l"(out + ro * Fout + f0), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7]) : "memory");
        } else {
            st4(out + ro * Fout + f0, make_float4(o[0], o[1], o[2], o[3]));
            st4(out + ro * Fout + f0 + 4, make_float4(o[4], o[5], o[6], o[7]));
        }
        if (argmax) *reinterpret_cast<uint2*>(argmax + ro * Fout + f0) = make_uint2(a0, a1);
    }
}

template <int Q>
__global__ void __launch_bounds__(256)
smallconv_bwd_fast_kernel(Planes P, RowMap rm, const float* __restrict__ bias, const float* __restrict__ out,
                          const uint8_t* __restrict__ argmax, const float* __restrict__ dout,
                          double* __restrict__ accg, long long Rout, int Fout)
{
    extern __shared__ float red[];                   // [(Q+2)*Fout]
    for (int i = threadIdx.x; i < (Q + 2) * Fout; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    const int Fg = Fout / 8;
    const int rows_per_block = blockDim.x / Fg;      // blockDim.x % Fg == 0
    const int f0 = (threadIdx.x % Fg) * 8, rl = threadIdx.x / Fg;
    float A[Q][8], s1[8], s2[8], bs[8];
    const float* cp[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        cp[q] = P.col(q);
#pragma unroll
        for (int v = 0; v < 8; ++v) A[q][v] = 0.f;
    }
#pragma unroll
    for (int v = 0; v < 8; ++v) { s1[v] = 0.f; s2[v] = 0.f; bs[v] = bias ? __ldg(bias + f0 + v) : 0.f; }
    const int Fin = P.Fin;
    const bool wide = (((uintptr_t)dout | (uintptr_t)out) & 31) == 0;
    for (long long t = (long long)blockIdx.x * rows_per_block + rl; t < Rout; t += (long long)gridDim.x * rows_per_block) {
        int b, mo;
        rm.decode(t, b, mo);
        const long long ro = (long long)b * rm.Mo + mo;
        float d[8], o[8];
        if (wide) {                                      // one 32-byte load per lane and tensor: half the wavefronts
            ldg8(dout + ro * Fout + f0, d);
            ldg8(out + ro * Fout + f0, o);
        } else {
            const float4 d0 = ldg4(dout + ro * Fout + f0), d1 = ldg4(dout + ro * Fout + f0 + 4);
            const float4 o0 = ldg4(out + ro * Fout + f0), o1 = ldg4(out + ro * Fout + f0 + 4);
            d[0] = d0.x; d[1] = d0.y; d[2] = d0.z; d[3] = d0.w; d[4] = d1.x; d[5] = d1.y; d[6] = d1.z; d[7] = d1.w;
            o[0] = o0.x; o[1] = o0.y; o[2] = o0.z; o[3] = o0.w; o[4] = o1.x; o[5] = o1.y; o[6] = o1.z; o[7] = o1.w;
        }
        const uint2 ab = __ldg(reinterpret_cast<const uint2*>(argmax + ro * Fout + f0));
        float x[4][Q];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long r = (long long)rm.xrow(b, mo, j) * Fin;
#pragma unroll
            for (int q = 0; q < Q; ++q) x[j][q] = __ldg(cp[q] + r);
        }
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            const float gz = o[v] > 0.f ? d[v] : 0.f;
            const uint32_t arg = ((v < 4 ? ab.x >> (8 * v) : ab.y >> (8 * (v - 4))) & 3u);
            s1[v] += gz;
            s2[v] = fmaf(gz, o[v] - bs[v], s2[v]);           // xhat = out - bias wherever gz != 0
            const bool b0 = (arg & 1u) != 0, b1 = (arg & 2u) != 0;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const float lo = b0 ? x[1][q] : x[0][q], hi = b0 ? x[3][q] : x[2][q];
                A[q][v] = fmaf(b1 ? hi : lo, gz, A[q][v]);
            }
        }
    }
    // lanes equal modulo Fg hold the same features: reduce inside the warp before touching shared memory
    const bool pw = Fg < 32 && (Fg & (Fg - 1)) == 0;
    const bool leader = !pw || (threadIdx.x & 31) < Fg;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const float a = pw ? warp_sum_mod(A[q][v], Fg) : A[q][v];
            if (leader) atomicAdd(&red[q * Fout + f0 + v], a);
        }
        const float a1 = pw ? warp_sum_mod(s1[v], Fg) : s1[v];
        const float a2 = pw ? warp_sum_mod(s2[v], Fg) : s2[v];
        if (leader) { atomicAdd(&red[Q * Fout + f0 + v], a1); atomicAdd(&red[(Q + 1) * Fout + f0 + v], a2); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (Q + 2) * Fout; i += blockDim.x) atomicAdd(accg + i, (double)red[i]);
}

// dW[fi*K+k, f] += rstd_f (A_qf - S1_f/n s_q - S2_f/n rstd_f (sum_q' G_qq' W_q'f - mean_f s_q));  dbias_f = S1_f
__global__ void smallconv_bwd_finalize_kernel(const double* __restrict__ acc, const double* __restrict__ gram,
                                              const float* __restrict__ W, const float* __restrict__ mean,
                                              const float* __restrict__ rstd, double count, int Fin, int K, int Fout,
                                              float* dW, float* dbias)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int Q = Fin * K;
    if (i >= Q * Fout) return;
    const int q = i / Fout, f = i - q * Fout;
    const int k = q / Fin, fi = q - k * Fin;
    const double S1 = acc[Q * Fout + f], S2 = acc[(Q + 1) * Fout + f];
    double g = acc[i];
    if (mean) {
        const double rs = (double)rstd[f], mu = (double)mean[f];
        double xy = 0.0;
        for (int b = 0; b < Q; ++b) {
            const int kb = b / Fin, fb = b - kb * Fin;
            xy += gram[Q + q * Q + b] * (double)W[(fb * K + kb) * Fout + f];
        }
        g = rs * (g - S1 / count * gram[q] - S2 / count * rs * (xy - mu * gram[q]));
    }
    dW[(fi * K + k) * Fout + f] += (float)g;
    if (q == 0 && dbias) dbias[f] = (float)S1;
}

// out[m][b][f] = x[b][m][f]: tiles of TM vertices staged through shared memory, coalesced on both sides.  Index arithmetic by
// shifts when B * F and F are powers of two (cosmo: 16, 1): the div / mod per element made this copy issue-bound (30 us for
// 34 MB of traffic).
template <bool POW2>
__global__ void __launch_bounds__(256)
interleave_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int M, int F, int TM, int lgBF, int lgF)
{
    extern __shared__ float tile[];                  // [B][TM*F + 1]
    const int m0 = blockIdx.x * TM;
    const int tm = min(TM, M - m0);
    const int w = tm * F, pitch = TM * F + 1;
    for (int b = 0; b < B; ++b) {
        const float* src = x + ((long long)b * M + m0) * F;
        for (int r = threadIdx.x; r < w; r += blockDim.x) tile[b * pitch + r] = __ldg(src + r);
    }
    __syncthreads();
    float* dst = out + (long long)m0 * B * F;
    const int BF = B * F;
    for (int i = threadIdx.x; i < B * w; i += blockDim.x) {
        int m, b, f;
        if (POW2) { m = i >> lgBF; const int r = i & (BF - 1); b = r >> lgF; f = r & (F - 1); }
        else { m = i / BF; const int r = i - m * BF; b = r / F; f = r - b * F; }
        dst[i] = tile[b * pitch + m * F + f];
    }
}

static inline bool smallconv_fast_ok(int Fout, int p, int pool, int act) {
    return Fout % 8 == 0 && 256 % (Fout / 8) == 0 && p == 4 && pool == DSB200_POOL_MAX && act == DSB200_ACT_RELU;
}

static inline bool smallconv_ok(int Fin, int K, int Fout) { return Fin * K >= 1 && Fin * K <= QMAX && Fout % 4 == 0 && Fout <= 256 && 256 % (Fout / 4) == 0; }

}  // namespace dsb200

using namespace dsb200;
#define STREAM ((cudaStream_t)stream)

template <int Q> static void launch_gram(Planes P, long long R, double* gram, int grid, cudaStream_t st)
{ gram_kernel<Q><<<grid, 256, 0, st>>>(P, R, gram); }
template <int Q> static void launch_fwd(Planes P, RowMap rm, const float* W, const float* mean, const float* rstd, const float* bias,
                                        float* out, long long Rout, int Fout, int p, int pool, int act, int grid, size_t smem, cudaStream_t st)
{ smallconv_fwd_kernel<Q><<<grid, 256, smem, st>>>(P, rm, W, mean, rstd, bias, out, Rout, Fout, p, pool, act); }
template <int Q> static void launch_fwd_fast(Planes P, RowMap rm, const float* W, const float* mean, const float* rstd,
                                             const float* bias, float* out, uint8_t* argmax, long long Rout, int Fout,
                                             int grid, size_t smem, cudaStream_t st)
{ smallconv_fwd_fast_kernel<Q><<<grid, 256, smem, st>>>(P, rm, W, mean, rstd, bias, out, argmax, Rout, Fout); }
template <int Q> static void launch_bwd_fast(Planes P, RowMap rm, const float* bias, const float* out, const uint8_t* argmax,
                                             const float* dout, double* acc, long long Rout, int Fout, int grid, size_t smem,
                                             cudaStream_t st)
{ smallconv_bwd_fast_kernel<Q><<<grid, 256, smem, st>>>(P, rm, bias, out, argmax, dout, acc, Rout, Fout); }
template <int Q> static void launch_bwd(Planes P, RowMap rm, const float* W, const float* mean, const float* rstd, const float* bias,
                                        const float* dout, double* acc, long long Rout, int Fout, int p, int pool, int act,
                                        int grid, size_t smem, cudaStream_t st)
{ smallconv_bwd_kernel<Q><<<grid, 256, smem, st>>>(P, rm, W, mean, rstd, bias, dout, acc, Rout, Fout, p, pool, act); }

#define DISPATCH_Q(Q, FN, ...) \
    switch (Q) { case 1: FN<1>(__VA_ARGS__); break; case 2: FN<2>(__VA_ARGS__); break; case 3: FN<3>(__VA_ARGS__); break; \
                 case 4: FN<4>(__VA_ARGS__); break; case 5: FN<5>(__VA_ARGS__); break; case 6: FN<6>(__VA_ARGS__); break; \
                 case 7: FN<7>(__VA_ARGS__); break; default: FN<8>(__VA_ARGS__); break; }

extern "C" int dsb200_smallconv_supported(int Fin, int K, int Fout) { return smallconv_ok(Fin, K, Fout) ? 1 : 0; }
extern "C" int dsb200_smallconv_has_argmax(int Fout, int p, int pool, int act) { return smallconv_fast_ok(Fout, p, pool, act) ? 1 : 0; }

extern "C" int dsb200_cheb_gram(const float* x0, const float* xk, long long R, int Fin, int K, double* gram, void* stream)
{
    DSB_REQUIRE(x0 && gram && R > 0 && (K == 1 || xk), "cheb_gram: bad arguments");
    DSB_REQUIRE(Fin * K <= QMAX, "cheb_gram: K*Fin too large");
    Planes P{x0, xk, R * (long long)Fin, Fin, K};
    const int grid = grid_for(R, 256, 4);
    DISPATCH_Q(Fin * K, launch_gram, P, R, gram, grid, STREAM);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_bn_sums_from_gram(const double* gram, const float* W, int Fin, int K, int Fout, double* sums, void* stream)
{
    DSB_REQUIRE(gram && W && sums && Fin * K <= QMAX && Fout > 0, "bn_sums_from_gram: bad arguments");
    bn_sums_from_gram_kernel<<<(Fout + 63) / 64, 64, 0, STREAM>>>(gram, W, Fin, K, Fout, sums);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_smallconv_fwd(const float* x0, const float* xk, const float* W, const float* mean, const float* rstd,
                                    const float* bias, float* out, int B, int M, int Fin, int K, int Fout, int p, int pool,
                                    int act, int interleaved, uint8_t* argmax, void* stream)
{
    DSB_REQUIRE(x0 && W && out && (K == 1 || xk) && B > 0 && M > 0 && p >= 1 && M % p == 0, "smallconv_fwd: bad arguments");
    DSB_REQUIRE(smallconv_ok(Fin, K, Fout), "smallconv_fwd: unsupported shape");
    DSB_REQUIRE((mean == nullptr) == (rstd == nullptr), "smallconv_fwd: mean and rstd go together");
    const long long R = (long long)B * M, Rout = R / p;
    DSB_REQUIRE(R < 0x7fffffffLL, "smallconv_fwd: too many rows");
    Planes P{x0, xk, R * (long long)Fin, Fin, K};
    const int Q = Fin * K;
    const size_t smem = (size_t)(Q + 3) * Fout * sizeof(float);
    const int grid = grid_for(Rout * (Fout / 4), 256, 8);
    RowMap rm{B, M, M / p, p, interleaved};
    if (smallconv_fast_ok(Fout, p, pool, act)) {
        const size_t smf = (size_t)(Q + 2) * Fout * sizeof(float);
        const int gf = grid_for(Rout * (Fout / 8), 256, 8);
        DISPATCH_Q(Q, launch_fwd_fast, P, rm, W, mean, rstd, bias, out, argmax, Rout, Fout, gf, smf, STREAM);
        DSB_LAUNCH_CHECK();
    }
    DSB_REQUIRE(!argmax, "smallconv_fwd: the arg-max record needs relu + max pooling over 4 vertices");
    DISPATCH_Q(Q, launch_fwd, P, rm, W, mean, rstd, bias, out, Rout, Fout, p, pool, act, grid, smem, STREAM);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_smallconv_bwd(const float* x0, const float* xk, const float* W, const float* mean, const float* rstd,
                                    const float* bias, const float* dout, double* acc, int B, int M, int Fin, int K, int Fout,
                                    int p, int pool, int act, int interleaved, const float* out, const uint8_t* argmax,
                                    void* stream)
{
    DSB_REQUIRE(x0 && W && dout && acc && (K == 1 || xk) && B > 0 && M > 0 && p >= 1 && M % p == 0, "smallconv_bwd: bad arguments");
    DSB_REQUIRE(smallconv_ok(Fin, K, Fout), "smallconv_bwd: unsupported shape");
    const long long R = (long long)B * M, Rout = R / p;
    DSB_REQUIRE(R < 0x7fffffffLL, "smallconv_bwd: too many rows");
    Planes P{x0, xk, R * (long long)Fin, Fin, K};
    const int Q = Fin * K;
    const size_t smem = (size_t)(2 * Q + 5) * Fout * sizeof(float);
    const int rpb = 256 / (Fout / 4);
    const int grid = grid_for((Rout + rpb - 1) / rpb * 256, 256, 4);
    RowMap rm{B, M, M / p, p, interleaved};
    if (out && argmax) {
        DSB_REQUIRE(smallconv_fast_ok(Fout, p, pool, act), "smallconv_bwd: arg-max path needs relu + max pooling over 4 vertices");
        const size_t smf = (size_t)(Q + 2) * Fout * sizeof(float);
        const int rpf = 256 / (Fout / 8);
        const int gf = grid_for((Rout + rpf - 1) / rpf * 256, 256, 4);
        DISPATCH_Q(Q, launch_bwd_fast, P, rm, bias, out, argmax, dout, acc, Rout, Fout, gf, smf, STREAM);
        DSB_LAUNCH_CHECK();
    }
    DISPATCH_Q(Q, launch_bwd, P, rm, W, mean, rstd, bias, dout, acc, Rout, Fout, p, pool, act, grid, smem, STREAM);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_smallconv_bwd_finalize(const double* acc, const double* gram, const float* W, const float* mean,
                                             const float* rstd, double count, int Fin, int K, int Fout, float* dW,
                                             float* dbias, void* stream)
{
    DSB_REQUIRE(acc && W && dW && Fin * K <= QMAX && Fout > 0, "smallconv_bwd_finalize: bad arguments");
    DSB_REQUIRE(!mean || (rstd && gram && count > 0), "smallconv_bwd_finalize: batch norm needs rstd, gram and count");
    const int n = Fin * K * Fout;
    smallconv_bwd_finalize_kernel<<<(n + 127) / 128, 128, 0, STREAM>>>(acc, gram, W, mean, rstd, count, Fin, K, Fout, dW, dbias);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_interleave(const float* x, float* out, int B, int M, int F, void* stream)
{
    DSB_REQUIRE(x && out && B > 0 && M > 0 && F > 0 && x != out, "interleave: bad arguments");
    int TM = 4096 / (B * F);                              // 16 KB tiles: several CTAs per SM
    if (TM < 1) TM = 1;
    if (TM > M) TM = M;
    const size_t smem = (size_t)B * (TM * F + 1) * sizeof(float);
    DSB_REQUIRE(smem <= 48 * 1024, "interleave: B*F too large");
    const int BF = B * F;
    int lgBF = 0, lgF = 0;
    while ((1 << lgBF) < BF) ++lgBF;
    while ((1 << lgF) < F) ++lgF;
    if ((1 << lgBF) == BF && (1 << lgF) == F)
        interleave_kernel<true><<<(M + TM - 1) / TM, 256, smem, STREAM>>>(x, out, B, M, F, TM, lgBF, lgF);
    else
        interleave_kernel<false><<<(M + TM - 1) / TM, 256, smem, STREAM>>>(x, out, B, M, F, TM, 0, 0);
    DSB_LAUNCH_CHECK();
}
