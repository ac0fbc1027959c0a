// K3 / K4 and the small operators around the ChebConv layers, all on [B, M, F] fp32 tensors:
//   batch-norm finalisation + moving statistics     (reference models.py:1194-1203, 840-842)
//   normalise -> bias -> activation -> pool, fwd/bwd  (models.py:1235-1243, 1122-1163)
//   vertex resize / unpool / concat                   (models.py:1138, 1157, 1347-1375, 1311)
//   statistical head                                  (models.py:1249-1270)
//   fc layers of the head                             (models.py:1205-1217)
//   softmax cross-entropy / MSE / L2 regulariser      (models.py:777-820)
//   optimizers with TF1 update rules                  (models.py:822-843)
#include "common.cuh"
#include "bf16.cuh"
#include <float.h>

namespace dsb200 {

// ------------------------------------------------------------------------------------------
// batch-norm statistics
__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, int F, float eps, float momentum,
                                   float* mean, float* rstd, float* moving_mean, float* moving_var)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const double mu = sums[f] / count;
    double var = sums[F + f] / count - mu * mu;            // biased variance (tf.nn.moments)
    if (var < 0.0) var = 0.0;
    mean[f] = (float)mu;
    rstd[f] = (float)(1.0 / sqrt(var + (double)eps));
    if (moving_mean) moving_mean[f] = moving_mean[f] * momentum + (float)mu * (1.f - momentum);
    if (moving_var) moving_var[f] = moving_var[f] * momentum + (float)var * (1.f - momentum);
}

__global__ void bn_from_moving_kernel(const float* __restrict__ mm, const float* __restrict__ mv, int F, float eps,
                                      float* mean, float* rstd)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    mean[f] = mm[f];
    rstd[f] = rsqrtf(mv[f] + eps);
}

__global__ void sums_to_float_kernel(const double* __restrict__ s, float* out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)s[i];
}

// ------------------------------------------------------------------------------------------
// normalise -> + bias -> activation -> pool over p consecutive vertices.
// One thread owns VEC features of one pooled (sample, vertex) row.
template <int VEC, int ACT, typename T>
__global__ void __launch_bounds__(256)
bn_act_pool_fwd_kernel(const T* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ rstd,
                       const float* __restrict__ bias, T* __restrict__ out,
                       long long Rout, int Fv, int p, int pool, int act)
{
    const long long total = Rout * Fv;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int F = Fv * VEC;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        const long long ro = g / Fv;
        const int f0 = (int)(g - ro * Fv) * VEC;
        float mu[VEC], rs[VEC], bs[VEC], acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            mu[v] = mean ? __ldg(mean + f0 + v) : 0.f;
            rs[v] = rstd ? __ldg(rstd + f0 + v) : 1.f;
            bs[v] = bias ? __ldg(bias + f0 + v) : 0.f;
            acc[v] = (pool == DSB200_POOL_MAX) ? -FLT_MAX : 0.f;
        }
        const T* yp = y + (ro * p) * F + f0;
        for (int j = 0; j < p; ++j) {
            float t[VEC];
            if (VEC == 4) { const float4 q = ldg4(yp + (long long)j * F); t[0] = q.x; t[1] = q.y; t[2] = q.z; t[3] = q.w; }
            else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) t[v] = ldg1(yp + (long long)j * F + v);
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float a = act_fwd_t<ACT>((t[v] - mu[v]) * rs[v] + bs[v], act);
                acc[v] = (pool == DSB200_POOL_MAX) ? fmaxf(acc[v], a) : acc[v] + a;
            }
        }
        if (pool != DSB200_POOL_MAX) {
            const float inv = 1.f / (float)p;
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[v] *= inv;
        }
        T* op = out + ro * F + f0;
        if (VEC == 4) st4(op, make_float4(acc[0], acc[1], acc[2], acc[3]));
        else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) st1(op + v, acc[v]);
        }
    }
}

// Lean forward for the combination every DeepSphere encoder layer uses (batch norm, relu, max pool over P children,
// F % 4 == 0), same arithmetic as the general kernel.  One thread = one float4 feature group of one pooled row, its P
// loads issued back to back; the per-feature parameters sit in shared memory.
template <int P, typename T>
__global__ void __launch_bounds__(256)
bn_relu_max_fwd_kernel(const T* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ rstd,
                       const float* __restrict__ bias, T* __restrict__ out, long long total, int Fv)
{
    extern __shared__ float par[];                          // [3][F]: mean, rstd, bias
    const int F = Fv * 4;
    for (int i = threadIdx.x; i < F; i += blockDim.x) {
        par[i] = __ldg(mean + i);
        par[F + i] = __ldg(rstd + i);
        par[2 * F + i] = bias ? __ldg(bias + i) : 0.f;
    }
    __syncthreads();
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    const long long ro = g / Fv;
    const int f0 = (int)(g - ro * Fv) * 4;
    const T* yp = y + (ro * P) * F + f0;
    float4 q[P];
#pragma unroll
    for (int j = 0; j < P; ++j) q[j] = ldg4(yp + (long long)j * F);
    const float4 mu = *reinterpret_cast<const float4*>(par + f0), rs = *reinterpret_cast<const float4*>(par + F + f0);
    const float4 bs = *reinterpret_cast<const float4*>(par + 2 * F + f0);
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);            // relu folded into the running maximum
#pragma unroll
    for (int j = 0; j < P; ++j) {
        r.x = fmaxf(r.x, (q[j].x - mu.x) * rs.x + bs.x); r.y = fmaxf(r.y, (q[j].y - mu.y) * rs.y + bs.y);
        r.z = fmaxf(r.z, (q[j].z - mu.z) * rs.z + bs.z); r.w = fmaxf(r.w, (q[j].w - mu.w) * rs.w + bs.w);
    }
    st4(out + ro * F + f0, r);
}

// Backward of the chain.  Threads are arranged [rows_per_block][Fv] so that a thread keeps the
// same feature group over its grid-stride loop and can accumulate sum(gz), sum(gz*xhat) in
// registers; one shared-memory and one double-precision global atomic pass at the end.
//   MODE 0: write gz and accumulate the sums          (no batch norm: gz is already dy)
//   MODE 1: accumulate the sums only                  (batch norm, pass 1: nothing is written)
//   MODE 2: write dy = rstd (gz - S1/n - xhat S2/n)   (batch norm, pass 2: gz is recomputed, never stored)
template <int VEC, typename T> __device__ __forceinline__ void load_vec(const T* p, float* v) {
    if (VEC == 4) { const float4 q = ldg4(p); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
    else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = ldg1(p + i);
    }
}
template <int VEC, typename T> __device__ __forceinline__ void store_vec(T* p, const float* v) {
    if (VEC == 4) st4(p, make_float4(v[0], v[1], v[2], v[3]));
    else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) st1(p + i, v[i]);
    }
}

template <int VEC, int MODE, int ACT, typename T>
__global__ void __launch_bounds__(256)
bn_act_pool_bwd_kernel(const T* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ rstd,
                       const float* __restrict__ bias, const T* __restrict__ dout, T* __restrict__ gz,
                       double* __restrict__ sums_bwd, double count, long long Rout, int Fv, int p, int pool, int act)
{
    extern __shared__ float red[];                          // [2][F]
    const int F = Fv * VEC;
    if (MODE != 2) {
        for (int i = threadIdx.x; i < 2 * F; i += blockDim.x) red[i] = 0.f;
        __syncthreads();
    }
    const int rows_per_block = blockDim.x / Fv;
    const int fvi = threadIdx.x % Fv, rl = threadIdx.x / Fv;
    const int f0 = fvi * VEC;
    float s1[VEC], s2[VEC], mu[VEC], rs[VEC], bs[VEC], m1[VEC], m2[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        s1[v] = 0.f; s2[v] = 0.f;
        mu[v] = mean ? __ldg(mean + f0 + v) : 0.f;
        rs[v] = rstd ? __ldg(rstd + f0 + v) : 1.f;
        bs[v] = bias ? __ldg(bias + f0 + v) : 0.f;
        m1[v] = MODE == 2 ? (float)(sums_bwd[f0 + v] / count) : 0.f;
        m2[v] = MODE == 2 ? (float)(sums_bwd[F + f0 + v] / count) : 0.f;
    }
    if (rl < rows_per_block) {
        constexpr int PC = 4;                               // sibling rows kept in registers
        for (long long ro = (long long)blockIdx.x * rows_per_block + rl; ro < Rout; ro += (long long)gridDim.x * rows_per_block) {
            float d[VEC];
            load_vec<VEC>(dout + ro * F + f0, d);
            const T* yp = y + (ro * p) * F + f0;
            T* gp = MODE != 1 ? gz + (ro * p) * F + f0 : nullptr;
            const bool is_max = pool == DSB200_POOL_MAX && p > 1;
            const float inv = is_max ? 1.f : 1.f / (float)p;
            if (p <= PC) {
                float yv[PC][VEC];
                int arg[VEC];
                float best[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) { best[v] = -FLT_MAX; arg[v] = 0; }
#pragma unroll
                for (int j = 0; j < PC; ++j) {
                    if (j < p) {
                        load_vec<VEC>(yp + (long long)j * F, yv[j]);
                        if (is_max) {
#pragma unroll
                            for (int v = 0; v < VEC; ++v) {
                                const float a = act_fwd_t<ACT>((yv[j][v] - mu[v]) * rs[v] + bs[v], act);
                                if (a > best[v]) { best[v] = a; arg[v] = j; }
                            }
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < PC; ++j) {
                    if (j < p) {
                        float o[VEC];
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            const float xh = (yv[j][v] - mu[v]) * rs[v];
                            float g = 0.f;
                            if (!is_max || arg[v] == j) g = d[v] * inv * act_grad_t<ACT>(xh + bs[v], act);
                            if (MODE != 2) { s1[v] += g; s2[v] = fmaf(g, xh, s2[v]); }
                            o[v] = MODE == 2 ? rs[v] * (g - m1[v] - xh * m2[v]) : g;
                        }
                        if (MODE != 1) store_vec<VEC>(gp + (long long)j * F, o);
                    }
                }
            } else {
                int arg[VEC];
                float best[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) { best[v] = -FLT_MAX; arg[v] = 0; }
                if (is_max) {
                    for (int j = 0; j < p; ++j) {
                        float yv[VEC];
                        load_vec<VEC>(yp + (long long)j * F, yv);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            const float a = act_fwd_t<ACT>((yv[v] - mu[v]) * rs[v] + bs[v], act);
                            if (a > best[v]) { best[v] = a; arg[v] = j; }
                        }
                    }
                }
                for (int j = 0; j < p; ++j) {
                    float yv[VEC], o[VEC];
                    load_vec<VEC>(yp + (long long)j * F, yv);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        const float xh = (yv[v] - mu[v]) * rs[v];
                        float g = 0.f;
                        if (!is_max || arg[v] == j) g = d[v] * inv * act_grad_t<ACT>(xh + bs[v], act);
                        if (MODE != 2) { s1[v] += g; s2[v] = fmaf(g, xh, s2[v]); }
                        o[v] = MODE == 2 ? rs[v] * (g - m1[v] - xh * m2[v]) : g;
                    }
                    if (MODE != 1) store_vec<VEC>(gp + (long long)j * F, o);
                }
            }
        }
    }
    if (MODE != 2) {
        // (all threads take part in the shuffles; idle threads carry zeros)
        const bool pw = Fv < 32 && (Fv & (Fv - 1)) == 0 && (blockDim.x % Fv) == 0;
        const bool leader = rl < rows_per_block && (!pw || (threadIdx.x & 31) < Fv);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const float a1 = pw ? warp_sum_mod(s1[v], Fv) : s1[v];
            const float a2 = pw ? warp_sum_mod(s2[v], Fv) : s2[v];
            if (leader) { atomicAdd(&red[f0 + v], a1); atomicAdd(&red[F + f0 + v], a2); }
        }
    }
    if (MODE != 2) {
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * F; i += blockDim.x) atomicAdd(sums_bwd + i, (double)red[i]);
    }
}

// Lean path of the two-pass batch-norm backward for the combination every DeepSphere encoder layer uses:
// relu + max pool over P consecutive vertices (or no pool, P = 1), F % 4 == 0.
//
// Pass 1 needs no pre-pool tensor at all.  gz is non-zero only at the arg-max child of a pooled row, and only
// where the pooled output is positive; there  xhat = out - bias  (out = relu(max_j xhat_j + bias)).  Hence
//   sum gz       = sum_{pooled rows} [out > 0] dout
//   sum gz*xhat  = sum_{pooled rows} [out > 0] dout (out - bias)
// read from the two POOLED tensors (2 * Y/P bytes instead of Y + Y/P).
template <typename T>
__global__ void __launch_bounds__(256)
bn_relu_pool_sums_kernel(const T* __restrict__ out, const T* __restrict__ dout, const float* __restrict__ bias,
                         double* __restrict__ sums_bwd, long long Rout, int Fv)
{
    extern __shared__ float red[];                          // [2][F]
    const int F = Fv * 4;
    for (int i = threadIdx.x; i < 2 * F; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    const int rows_per_block = blockDim.x / Fv;
    const int fvi = threadIdx.x % Fv, rl = threadIdx.x / Fv;
    const int f0 = fvi * 4;
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    float bs[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) bs[v] = bias ? __ldg(bias + f0 + v) : 0.f;
    if (rl < rows_per_block) {
        const long long step = (long long)gridDim.x * rows_per_block;
#pragma unroll 4
        for (long long ro = (long long)blockIdx.x * rows_per_block + rl; ro < Rout; ro += step) {
            const float4 o = ldg4(out + ro * F + f0);
            const float4 d = ldg4(dout + ro * F + f0);
            const float ov[4] = {o.x, o.y, o.z, o.w}, dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const float g = ov[v] > 0.f ? dv[v] : 0.f;
                s1[v] += g;
                s2[v] = fmaf(g, ov[v] - bs[v], s2[v]);
            }
        }
    }
    const bool pw = Fv < 32 && (Fv & (Fv - 1)) == 0 && (blockDim.x % Fv) == 0;
    const bool leader = rl < rows_per_block && (!pw || (threadIdx.x & 31) < Fv);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const float a1 = pw ? warp_sum_mod(s1[v], Fv) : s1[v];
        const float a2 = pw ? warp_sum_mod(s2[v], Fv) : s2[v];
        if (leader) { atomicAdd(&red[f0 + v], a1); atomicAdd(&red[F + f0 + v], a2); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * F; i += blockDim.x) atomicAdd(sums_bwd + i, (double)red[i]);
}

// Pass 2: dy = rstd (gz - S1/n - xhat S2/n); one thread = one float4 feature group of one pooled row and its
// P children.  The arg-max is recomputed from y (first maximum wins, as in the forward pass).
template <int P, typename T>
__global__ void __launch_bounds__(256)
bn_relu_max_bwd_dy_kernel(const T* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ rstd,
                          const float* __restrict__ bias, const T* __restrict__ dout, T* __restrict__ dy,
                          const double* __restrict__ sums_bwd, double count, long long total, int Fv)
{
    extern __shared__ float par[];                          // [5][F]: mean, rstd, bias, S1/n, S2/n
    const int F = Fv * 4;
    for (int i = threadIdx.x; i < F; i += blockDim.x) {
        par[i] = __ldg(mean + i);
        par[F + i] = __ldg(rstd + i);
        par[2 * F + i] = bias ? __ldg(bias + i) : 0.f;
        par[3 * F + i] = (float)(sums_bwd[i] / count);
        par[4 * F + i] = (float)(sums_bwd[F + i] / count);
    }
    __syncthreads();
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    const long long ro = g / Fv;
    const int f0 = (int)(g - ro * Fv) * 4;
    const float4 mu4 = *reinterpret_cast<const float4*>(par + f0), rs4 = *reinterpret_cast<const float4*>(par + F + f0);
    const float4 bs4 = *reinterpret_cast<const float4*>(par + 2 * F + f0);
    const float4 a4 = *reinterpret_cast<const float4*>(par + 3 * F + f0), b4 = *reinterpret_cast<const float4*>(par + 4 * F + f0);
    const float mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w}, rs[4] = {rs4.x, rs4.y, rs4.z, rs4.w}, bs[4] = {bs4.x, bs4.y, bs4.z, bs4.w};
    const float m1[4] = {a4.x, a4.y, a4.z, a4.w}, m2[4] = {b4.x, b4.y, b4.z, b4.w};
    const float4 d4 = ldg4(dout + ro * F + f0);
    const float d[4] = {d4.x, d4.y, d4.z, d4.w};
    const T* yp = y + (ro * P) * F + f0;
    T* op = dy + (ro * P) * F + f0;
    float xh[P][4];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const float4 q = ldg4(yp + (long long)j * F);
        xh[j][0] = (q.x - mu[0]) * rs[0]; xh[j][1] = (q.y - mu[1]) * rs[1];
        xh[j][2] = (q.z - mu[2]) * rs[2]; xh[j][3] = (q.w - mu[3]) * rs[3];
    }
    int arg[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        float best = xh[0][v] + bs[v];
        arg[v] = 0;
#pragma unroll
        for (int j = 1; j < P; ++j) {
            const float t = xh[j][v] + bs[v];
            if (t > best) { best = t; arg[v] = j; }
        }
        if (!(best > 0.f)) arg[v] = -1;                      // relu inactive: no gradient at all
    }
#pragma unroll
    for (int j = 0; j < P; ++j) {
        float o[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const float gg = arg[v] == j ? d[v] : 0.f;
            o[v] = rs[v] * (gg - m1[v] - xh[j][v] * m2[v]);
        }
        st4(op + (long long)j * F, make_float4(o[0], o[1], o[2], o[3]));
    }
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const double* __restrict__ sums_bwd, double count, float* gz, long long R, int F)
{
    const long long total = R * F;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        const int f = (int)(g % F);
        const float rs = __ldg(rstd + f);
        const float xh = (__ldg(y + g) - __ldg(mean + f)) * rs;
        const float m1 = (float)(sums_bwd[f] / count), m2 = (float)(sums_bwd[F + f] / count);
        gz[g] = rs * (gz[g] - m1 - xh * m2);
    }
}

// ------------------------------------------------------------------------------------------
// vertex resize / unpool / concat
// rows of a [M, F] matrix picked by index: out[i, :] = x[idx[i], :]  (send buffer of the halo exchange)
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx, long long n, int F, float* __restrict__ out)
{
    const long long total = n * F;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        const long long i = g / F;
        const int f = (int)(g - i * F);
        out[g] = __ldg(x + (long long)__ldg(idx + i) * F + f);
    }
}

__global__ void __launch_bounds__(256)
vertex_resize_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int Min, int Mout, int F, float fill)
{
    const long long total = (long long)B * Mout * F;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long per = (long long)Mout * F;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        const long long b = g / per;
        const long long rem = g - b * per;                  // m * F + f
        out[g] = (rem < (long long)Min * F) ? __ldg(x + b * (long long)Min * F + rem) : fill;
    }
}

__global__ void __launch_bounds__(256)
unpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, long long Rin, int F, int p, int mode)
{
    const long long total = Rin * p * F;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        const long long ro = g / F;
        const int f = (int)(g - ro * F);
        const long long ri = ro / p;
        const int j = (int)(ro - ri * p);
        const float v = __ldg(x + ri * F + f);
        out[g] = (mode == DSB200_UNPOOL_REPEAT || j == 0) ? v : 0.f;
    }
}

__global__ void __launch_bounds__(256)
unpool_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dx, long long Rin, int F, int p, int mode)
{
    const long long total = Rin * F;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        const long long ri = g / F;
        const int f = (int)(g - ri * F);
        const float* dp = dout + (ri * p) * F + f;
        float s = __ldg(dp);
        if (mode == DSB200_UNPOOL_REPEAT)
            for (int j = 1; j < p; ++j) s += __ldg(dp + (long long)j * F);
        dx[g] = s;
    }
}

__global__ void __launch_bounds__(256)
concat_f_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                long long R, int Fa, int Fb)
{
    const int F = Fa + Fb;
    const long long total = R * F;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        const long long r = g / F;
        const int f = (int)(g - r * F);
        out[g] = f < Fa ? __ldg(a + r * Fa + f) : __ldg(b + r * Fb + (f - Fa));
    }
}

__global__ void __launch_bounds__(256)
split_f_kernel(const float* __restrict__ dout, float* __restrict__ da, float* db, long long R, int Fa, int Fb, int acc_b)
{
    const int F = Fa + Fb;
    const long long total = R * F;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        const long long r = g / F;
        const int f = (int)(g - r * F);
        const float v = __ldg(dout + g);
        if (f < Fa) da[r * Fa + f] = v;
        else { float* q = db + r * Fb + (f - Fa); *q = acc_b ? *q + v : v; }
    }
}

// ------------------------------------------------------------------------------------------
// statistical head: one warp per (sample, feature); lanes stride over the vertices.
__global__ void __launch_bounds__(256)
stat_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int M, int F, int kind)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B * F) return;
    const int b = warp / F, f = warp - b * F;
    const float* xp = x + (long long)b * M * F + f;
    const int Fo = (kind == DSB200_STAT_MEANVAR) ? 2 * F : F;
    if (kind == DSB200_STAT_MAX) {
        float mx = -FLT_MAX;
        for (int m = lane; m < M; m += 32) mx = fmaxf(mx, __ldg(xp + (long long)m * F));
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) out[(long long)b * Fo + f] = mx;
        return;
    }
    float s = 0.f;
    for (int m = lane; m < M; m += 32) s += __ldg(xp + (long long)m * F);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)M;
    if (kind == DSB200_STAT_MEAN) { if (lane == 0) out[(long long)b * Fo + f] = mean; return; }
    float v = 0.f;
    for (int m = lane; m < M; m += 32) { const float d = __ldg(xp + (long long)m * F) - mean; v = fmaf(d, d, v); }
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    v /= (float)M;
    if (lane == 0) {
        if (kind == DSB200_STAT_VAR) out[(long long)b * Fo + f] = v;
        else { out[(long long)b * Fo + f] = mean; out[(long long)b * Fo + F + f] = v; }
    }
}

__global__ void __launch_bounds__(256)
stat_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dout, float* __restrict__ dx,
                int B, int M, int F, int kind)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B * F) return;
    const int b = warp / F, f = warp - b * F;
    const float* xp = x + (long long)b * M * F + f;
    float* dp = dx + (long long)b * M * F + f;
    const int Fo = (kind == DSB200_STAT_MEANVAR) ? 2 * F : F;
    const float invM = 1.f / (float)M;
    if (kind == DSB200_STAT_MAX) {
        float mx = -FLT_MAX; int am = 0x7fffffff;
        for (int m = lane; m < M; m += 32) { const float v = __ldg(xp + (long long)m * F); if (v > mx) { mx = v; am = m; } }
        for (int o = 16; o; o >>= 1) {
            const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oam = __shfl_xor_sync(0xffffffffu, am, o);
            if (omx > mx || (omx == mx && oam < am)) { mx = omx; am = oam; }
        }
        const float g = __ldg(dout + (long long)b * Fo + f);
        for (int m = lane; m < M; m += 32) dp[(long long)m * F] = (m == am) ? g : 0.f;
        return;
    }
    if (kind == DSB200_STAT_MEAN) {
        const float g = __ldg(dout + (long long)b * Fo + f) * invM;
        for (int m = lane; m < M; m += 32) dp[(long long)m * F] = g;
        return;
    }
    float s = 0.f;
    for (int m = lane; m < M; m += 32) s += __ldg(xp + (long long)m * F);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * invM;
    const float gm = (kind == DSB200_STAT_MEANVAR) ? __ldg(dout + (long long)b * Fo + f) * invM : 0.f;
    const float gv = __ldg(dout + (long long)b * Fo + (kind == DSB200_STAT_MEANVAR ? F : 0) + f) * 2.f * invM;
    for (int m = lane; m < M; m += 32) dp[(long long)m * F] = gm + gv * (__ldg(xp + (long long)m * F) - mean);
}

// ------------------------------------------------------------------------------------------
// fc layers of the head: tiny matrices (N = batch <= a few hundred rows) -> one thread per output.
__global__ void __launch_bounds__(256)
fc_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b,
              float* __restrict__ y, int N, int Min, int Mout, int act)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N * Mout) return;
    const int n = g / Mout, o = g - n * Mout;
    float acc = b ? __ldg(b + o) : 0.f;
    const float* xp = x + (long long)n * Min;
    for (int i = 0; i < Min; ++i) acc = fmaf(__ldg(xp + i), __ldg(W + (long long)i * Mout + o), acc);
    y[g] = act_fwd(acc, act);
}

__device__ __forceinline__ float fc_dz(float yv, float dyv, int act)
{
    // derivative of the activation expressed through its OUTPUT y (the fc layers keep y only)
    switch (act) {
        case DSB200_ACT_RELU: return yv > 0.f ? dyv : 0.f;
        case DSB200_ACT_ELU: return yv > 0.f ? dyv : dyv * (yv + 1.f);                 // exp(z) = y + 1 for z <= 0
        case DSB200_ACT_TANH: return dyv * (1.f - yv * yv);
        case DSB200_ACT_SIGMOID: return dyv * yv * (1.f - yv);
        case DSB200_ACT_LEAKY_RELU: return yv > 0.f ? dyv : 0.2f * dyv;
        case DSB200_ACT_SOFTPLUS: return dyv * (1.f - expf(-yv));                      // sigmoid(z) = 1 - exp(-softplus(z))
        case DSB200_ACT_RELU6: return (yv > 0.f && yv < 6.f) ? dyv : 0.f;
        default: return dyv;
    }
}

__global__ void __launch_bounds__(256)
fc_bwd_dx_kernel(const float* __restrict__ W, const float* __restrict__ y, const float* __restrict__ dy,
                 float* __restrict__ dx, int N, int Min, int Mout, int act)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N * Min) return;
    const int n = g / Min, i = g - n * Min;
    float acc = 0.f;
    for (int o = 0; o < Mout; ++o)
        acc = fmaf(fc_dz(__ldg(y + (long long)n * Mout + o), __ldg(dy + (long long)n * Mout + o), act),
                   __ldg(W + (long long)i * Mout + o), acc);
    dx[g] = acc;
}

__global__ void __launch_bounds__(256)
fc_bwd_dw_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy,
                 float* dW, float* db, int N, int Min, int Mout, int act)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (Min + 1) * Mout) return;
    const int i = g / Mout, o = g - i * Mout;              // i == Min -> bias row
    float acc = 0.f;
    for (int n = 0; n < N; ++n) {
        const float dz = fc_dz(__ldg(y + (long long)n * Mout + o), __ldg(dy + (long long)n * Mout + o), act);
        acc = fmaf(i < Min ? __ldg(x + (long long)n * Min + i) : 1.f, dz, acc);
    }
    if (i < Min) dW[g] += acc;
    else if (db) db[o] += acc;
}

// ------------------------------------------------------------------------------------------
// losses
__device__ __forceinline__ float block_sum(float v, float* sh)
{
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = v;
    __syncthreads();
    v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
    if (w == 0) for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;                                               // valid in thread 0
}

// ---- learned histogram (cgcnn.learned_histogram, models.py:1166-1192; the `statistics='histogram'` head):
//   hist[b, f, j] = scale * mean_m relu(1 - |x[b, m, f] - c[f, j]| * |w[f, j]|),   scale = bins / initial_range / 4
// Block = (32 features) x (8 vertex lanes) of one sample: lanes read consecutive features (coalesced), every thread keeps
// its `bins` partial sums in registers, the 8 vertex lanes are reduced through shared memory.
constexpr int kHistBins = 32;                    // register budget: bins <= 32 (the reference hard-codes 20)

__global__ void __launch_bounds__(256)
hist_fwd_kernel(const float* __restrict__ x, const float* __restrict__ cen, const float* __restrict__ wid,
                float* __restrict__ out, int M, int F, int bins, float scale_over_M)
{
    __shared__ float red[8][32][kHistBins + 1];
    const int fl = threadIdx.x & 31, ml = threadIdx.x >> 5;
    const int f = blockIdx.x * 32 + fl, b = blockIdx.y;
    float acc[kHistBins], c[kHistBins], w[kHistBins];
#pragma unroll
    for (int j = 0; j < kHistBins; ++j) {
        acc[j] = 0.f;
        c[j] = (f < F && j < bins) ? __ldg(cen + f * bins + j) : 0.f;
        w[j] = (f < F && j < bins) ? fabsf(__ldg(wid + f * bins + j)) : 0.f;
    }
    if (f < F) {
        const float* xb = x + (long long)b * M * F + f;
        for (int m = ml; m < M; m += 8) {
            const float v = __ldg(xb + (long long)m * F);
#pragma unroll
            for (int j = 0; j < kHistBins; ++j) acc[j] += fmaxf(1.f - fabsf(v - c[j]) * w[j], 0.f);
        }
    }
#pragma unroll
    for (int j = 0; j < kHistBins; ++j) red[ml][fl][j] = acc[j];
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * bins; i += 256) {
        const int ff = i / bins, j = i - ff * bins;
        if (blockIdx.x * 32 + ff < F) {
            float t = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) t += red[q][ff][j];
            out[((long long)b * F + blockIdx.x * 32 + ff) * bins + j] = t * scale_over_M;
        }
    }
}

// dx[b, m, f] = sum_j dh[b, f, j] * s * (active ? -sign(x - c) |w| : 0);  dc += dh s (active ? sign(x - c) |w| : 0);
// dw += dh s (active ? -|x - c| sign(w) : 0);  active = 1 - |x - c| |w| > 0;  s = scale / M   (autodiff of 1187-1191)
__global__ void __launch_bounds__(256)
hist_bwd_kernel(const float* __restrict__ x, const float* __restrict__ cen, const float* __restrict__ wid,
                const float* __restrict__ dh, float* __restrict__ dx, float* dcen, float* dwid,
                int M, int F, int bins, float scale_over_M)
{
    __shared__ float red[8][32][kHistBins + 1];
    const int fl = threadIdx.x & 31, ml = threadIdx.x >> 5;
    const int f = blockIdx.x * 32 + fl, b = blockIdx.y;
    float c[kHistBins], w[kHistBins], sw[kHistBins], g[kHistBins], dc[kHistBins], dw[kHistBins];
#pragma unroll
    for (int j = 0; j < kHistBins; ++j) {
        const bool on = f < F && j < bins;
        c[j] = on ? __ldg(cen + f * bins + j) : 0.f;
        const float wr = on ? __ldg(wid + f * bins + j) : 0.f;
        w[j] = fabsf(wr);
        sw[j] = wr > 0.f ? 1.f : (wr < 0.f ? -1.f : 0.f);
        g[j] = on ? __ldg(dh + ((long long)b * F + f) * bins + j) * scale_over_M : 0.f;
        dc[j] = 0.f; dw[j] = 0.f;
    }
    if (f < F) {
        const long long base = (long long)b * M * F + f;
        for (int m = ml; m < M; m += 8) {
            const float v = __ldg(x + base + (long long)m * F);
            float d = 0.f;
#pragma unroll
            for (int j = 0; j < kHistBins; ++j) {
                const float diff = v - c[j], ad = fabsf(diff);
                const float sg = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
                const float ga = (1.f - ad * w[j] > 0.f) ? g[j] : 0.f;
                d = fmaf(-sg * w[j], ga, d);
                dc[j] = fmaf(sg * w[j], ga, dc[j]);
                dw[j] = fmaf(-ad * sw[j], ga, dw[j]);
            }
            dx[base + (long long)m * F] = d;
        }
    }
    for (int pass = 0; pass < 2; ++pass) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kHistBins; ++j) red[ml][fl][j] = pass ? dw[j] : dc[j];
        __syncthreads();
        float* dst = pass ? dwid : dcen;
        for (int i = threadIdx.x; i < 32 * bins; i += 256) {
            const int ff = i / bins, j = i - ff * bins;
            if (blockIdx.x * 32 + ff < F) {
                float t = 0.f;
#pragma unroll
                for (int q = 0; q < 8; ++q) t += red[q][ff][j];
                atomicAdd(dst + (blockIdx.x * 32 + ff) * bins + j, t);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
softmax_ce_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels, float* __restrict__ dlogits,
                  float* loss_out, long long R, int C, float gscale, const float* __restrict__ cw)
{
    __shared__ float sh[8];
    float local = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float invR = 1.f / (float)R;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < R; r += stride) {
        const float* lp = logits + r * C;
        float mx = -FLT_MAX;
        for (int c = 0; c < C; ++c) mx = fmaxf(mx, __ldg(lp + c));
        float se = 0.f;
        for (int c = 0; c < C; ++c) se += expf(__ldg(lp + c) - mx);
        const float lse = mx + logf(se);
        const int lab = __ldg(labels + r);
        const float wr = cw ? __ldg(cw + lab) : 1.f;                     // class weight of the row (models.py:795-802)
        local = fmaf(wr, lse - __ldg(lp + lab), local);
        if (dlogits) {
            float* dp = dlogits + r * C;
            const float k = gscale * invR * wr;
            for (int c = 0; c < C; ++c) dp[c] = (expf(__ldg(lp + c) - lse) - (c == lab ? 1.f : 0.f)) * k;
        }
    }
    const float tot = block_sum(local, sh);
    if (threadIdx.x == 0) atomicAdd(loss_out, tot * invR);
}

__global__ void __launch_bounds__(256)
mse_kernel(const float* __restrict__ pred, const float* __restrict__ target, float* __restrict__ dpred,
           float* loss_out, long long n, float gscale)
{
    __shared__ float sh[8];
    float local = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float invn = 1.f / (float)n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float d = __ldg(pred + i) - __ldg(target + i);
        local = fmaf(d, d, local);
        if (dpred) dpred[i] = 2.f * d * invn * gscale;
    }
    const float tot = block_sum(local, sh);
    if (threadIdx.x == 0) atomicAdd(loss_out, tot * invn);
}

__global__ void __launch_bounds__(256)
softmax_argmax_kernel(const float* __restrict__ logits, float* __restrict__ probs, int32_t* __restrict__ pred,
                      long long R, int C)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < R; r += stride) {
        const float* lp = logits + r * C;
        float mx = -FLT_MAX; int am = 0;
        for (int c = 0; c < C; ++c) { const float v = __ldg(lp + c); if (v > mx) { mx = v; am = c; } }
        if (pred) pred[r] = am;
        if (probs) {
            float se = 0.f;
            for (int c = 0; c < C; ++c) se += expf(__ldg(lp + c) - mx);
            const float inv = 1.f / se;
            for (int c = 0; c < C; ++c) probs[r * C + c] = expf(__ldg(lp + c) - mx) * inv;
        }
    }
}

__global__ void __launch_bounds__(256)
l2_reg_kernel(const float* __restrict__ w, const float* __restrict__ coef, float* g, long long n, float* loss_out)
{
    __shared__ float sh[8];
    float local = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float c = __ldg(coef + i), v = __ldg(w + i);
        local = fmaf(0.5f * c * v, v, local);
        if (g) g[i] = fmaf(c, v, g[i]);
    }
    const float tot = block_sum(local, sh);
    if (threadIdx.x == 0 && loss_out) atomicAdd(loss_out, tot);
}

// ------------------------------------------------------------------------------------------
// optimizers (TF 1.10 update rules; epsilon of Adam is outside the bias correction)
__global__ void __launch_bounds__(256)
optimizer_kernel(int kind, float* w, const float* __restrict__ g, float* s1, float* s2,
                 const float* __restrict__ hyper, long long n)
{
    const float lr = hyper[0], h1 = hyper[1], h2 = hyper[2], eps = hyper[3], lr_t = hyper[4], gs = hyper[5];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gi = __ldg(g + i) * gs;
        float wi = w[i];
        if (kind == DSB200_OPT_SGD) {
            wi -= lr * gi;
        } else if (kind == DSB200_OPT_MOMENTUM) {
            const float a = h1 * s1[i] + gi;                // accum = momentum * accum + grad
            s1[i] = a;
            wi -= lr * a;
        } else if (kind == DSB200_OPT_ADAM) {
            const float m = h1 * s1[i] + (1.f - h1) * gi;
            const float v = h2 * s2[i] + (1.f - h2) * gi * gi;
            s1[i] = m; s2[i] = v;
            wi -= lr_t * m / (sqrtf(v) + eps);
        } else {                                            // RMSProp: ms, mom
            const float ms = h1 * s1[i] + (1.f - h1) * gi * gi;
            const float mom = h2 * s2[i] + lr * gi / sqrtf(ms + eps);
            s1[i] = ms; s2[i] = mom;
            wi -= mom;
        }
        w[i] = wi;
    }
}

__global__ void __launch_bounds__(256) fill_kernel(float* x, float v, long long n)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = v;
}

__global__ void __launch_bounds__(256) axpy_kernel(const float* __restrict__ x, float* y, float a, long long n)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = fmaf(a, __ldg(x + i), y[i]);
}

// out[i] = a[i] * x[i * F + ch]: one feature channel of the input batch as a per-vertex mask (masked regression loss,
// reference deepsphere/models.py:785-787: `predictions * data[..., -2]`, `labels * data[..., -1]`)
__global__ void __launch_bounds__(256) mul_channel_kernel(const float* __restrict__ a, const float* __restrict__ x, int F, int ch,
                                                          float* __restrict__ out, long long n)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = a[i] * __ldg(x + i * F + ch);
}

__global__ void __launch_bounds__(256) scale_kernel(float* x, float a, long long n)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] *= a;
}

__global__ void __launch_bounds__(256)
dropout_kernel(const float* __restrict__ x, const float* __restrict__ u, float keep, float* __restrict__ out, long long n)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float inv = 1.f / keep;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (__ldg(u + i) < keep) ? __ldg(x + i) * inv : 0.f;
}

// storage-type conversion of an activation tensor (BF16 mode: operators without a bf16 kernel run in fp32 between two casts)
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) cast_kernel(const TI* __restrict__ x, TO* __restrict__ out, long long n4, long long n)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) st4(out + 4 * i, ldg4(x + 4 * i));
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - 4 * n4)) st1(out + 4 * n4 + threadIdx.x, ldg1(x + 4 * n4 + threadIdx.x));
}

static inline bool aligned16(const void* a, const void* b = nullptr, const void* c = nullptr)
{
    return (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0;
}

}  // namespace dsb200

using namespace dsb200;
#define STREAM ((cudaStream_t)stream)

extern "C" int dsb200_bn_finalize(const double* sums, double count, int F, float eps, float momentum,
                                  float* mean, float* rstd, float* moving_mean, float* moving_var, void* stream)
{
    DSB_REQUIRE(sums && mean && rstd && F > 0 && count > 0, "bn_finalize: bad arguments");
    bn_finalize_kernel<<<(F + 127) / 128, 128, 0, STREAM>>>(sums, count, F, eps, momentum, mean, rstd, moving_mean, moving_var);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_bn_from_moving(const float* moving_mean, const float* moving_var, int F, float eps,
                                     float* mean, float* rstd, void* stream)
{
    DSB_REQUIRE(moving_mean && moving_var && mean && rstd && F > 0, "bn_from_moving: bad arguments");
    bn_from_moving_kernel<<<(F + 127) / 128, 128, 0, STREAM>>>(moving_mean, moving_var, F, eps, mean, rstd);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_sums_to_float(const double* sums, float* out, int n, void* stream)
{
    DSB_REQUIRE(sums && out && n > 0, "sums_to_float: bad arguments");
    sums_to_float_kernel<<<(n + 127) / 128, 128, 0, STREAM>>>(sums, out, n);
    DSB_LAUNCH_CHECK();
}

template <typename T>
static int bn_act_pool_fwd_launch(const T* y, const float* mean, const float* rstd, const float* bias,
                                  T* out, int B, int M, int F, int p, int pool, int act, void* stream)
{
    DSB_REQUIRE(y && out && B > 0 && M > 0 && F > 0 && p >= 1, "bn_act_pool_fwd: bad arguments");
    DSB_REQUIRE(M % p == 0, "bn_act_pool_fwd: M must be a multiple of p");
    DSB_REQUIRE((mean == nullptr) == (rstd == nullptr), "bn_act_pool_fwd: mean and rstd go together");
    const long long Rout = (long long)B * (M / p);
    if (F % 4 == 0 && F <= 2048 && aligned16(y, out) && act == DSB200_ACT_RELU && mean && rstd &&
        (p == 1 || (pool == DSB200_POOL_MAX && (p == 2 || p == 4)))) {
        const long long total = Rout * (F / 4);
        const long long blocks = (total + 255) / 256;
        if (blocks <= 0x7fffffffLL) {
            const size_t sm = 3 * (size_t)F * sizeof(float);
#define DSB_LEAN(PP) bn_relu_max_fwd_kernel<PP, T><<<(unsigned)blocks, 256, sm, STREAM>>>(y, mean, rstd, bias, out, total, F / 4)
            if (p == 1) DSB_LEAN(1); else if (p == 2) DSB_LEAN(2); else DSB_LEAN(4);
#undef DSB_LEAN
            DSB_LAUNCH_CHECK();
        }
    }
#define DSB_FWD(V, A, G) bn_act_pool_fwd_kernel<V, A, T><<<G, 256, 0, STREAM>>>(y, mean, rstd, bias, out, Rout, F / V, p, pool, act)
    if (F % 4 == 0 && aligned16(y, out)) {
        const int g = grid_for(Rout * (F / 4), 256);
        if (act == DSB200_ACT_RELU) DSB_FWD(4, DSB200_ACT_RELU, g);
        else if (act == DSB200_ACT_IDENTITY) DSB_FWD(4, DSB200_ACT_IDENTITY, g);
        else DSB_FWD(4, -1, g);
    } else {
        const int g = grid_for(Rout * F, 256);
        DSB_FWD(1, -1, g);
    }
#undef DSB_FWD
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_bn_act_pool_fwd(const float* y, const float* mean, const float* rstd, const float* bias,
                                      float* out, int B, int M, int F, int p, int pool, int act, void* stream)
{
    return bn_act_pool_fwd_launch<float>(y, mean, rstd, bias, out, B, M, F, p, pool, act, stream);
}

extern "C" int dsb200_bn_act_pool_fwd_bf16(const void* y, const float* mean, const float* rstd, const float* bias,
                                           void* out, int B, int M, int F, int p, int pool, int act, void* stream)
{
    return bn_act_pool_fwd_launch<bf16>((const bf16*)y, mean, rstd, bias, (bf16*)out, B, M, F, p, pool, act, stream);
}

template <int MODE, typename T>
static int bn_act_pool_bwd_launch(const T* y, const float* mean, const float* rstd, const float* bias,
                                  const T* dout, T* gz, double* sums_bwd, double count,
                                  int B, int M, int F, int p, int pool, int act, cudaStream_t st)
{
    DSB_REQUIRE(y && dout && sums_bwd && B > 0 && M > 0 && F > 0 && p >= 1, "bn_act_pool_bwd: bad arguments");
    DSB_REQUIRE(M % p == 0, "bn_act_pool_bwd: M must be a multiple of p");
    DSB_REQUIRE(F <= 4096, "bn_act_pool_bwd: F too large");
    const long long Rout = (long long)B * (M / p);
    const size_t smem = 2 * (size_t)F * sizeof(float);
    const bool a16 = ((((uintptr_t)y | (uintptr_t)dout | (uintptr_t)gz) & 15) == 0);
    if (MODE == 2 && F % 4 == 0 && F <= 2048 && a16 && act == DSB200_ACT_RELU && mean && rstd &&
        (p == 1 || (pool == DSB200_POOL_MAX && (p == 2 || p == 4)))) {
        const long long total = Rout * (F / 4);
        const long long blocks = (total + 255) / 256;
        if (blocks <= 0x7fffffffLL) {
            const size_t sm = 5 * (size_t)F * sizeof(float);
#define DSB_DY(PP) bn_relu_max_bwd_dy_kernel<PP, T><<<(unsigned)blocks, 256, sm, st>>>(y, mean, rstd, bias, dout, gz, sums_bwd, count, total, F / 4)
            if (p == 1) DSB_DY(1); else if (p == 2) DSB_DY(2); else DSB_DY(4);
#undef DSB_DY
            DSB_LAUNCH_CHECK();
        }
    }
#define DSB_BWD(V, A) bn_act_pool_bwd_kernel<V, MODE, A, T><<<grid, 256, smem, st>>>(y, mean, rstd, bias, dout, gz, sums_bwd, count, Rout, F / V, p, pool, act)
    if (F % 4 == 0 && F / 4 <= 256 && a16) {
        const int Fv = F / 4, rpb = 256 / Fv;
        const int grid = grid_for((Rout + rpb - 1) / rpb * 256, 256, 8);
        if (act == DSB200_ACT_RELU) DSB_BWD(4, DSB200_ACT_RELU);
        else if (act == DSB200_ACT_IDENTITY) DSB_BWD(4, DSB200_ACT_IDENTITY);
        else DSB_BWD(4, -1);
    } else {
        DSB_REQUIRE(F <= 256, "bn_act_pool_bwd: F must be a multiple of 4 when > 256");
        const int rpb = 256 / F;
        const int grid = grid_for((Rout + rpb - 1) / rpb * 256, 256, 8);
        DSB_BWD(1, -1);
    }
#undef DSB_BWD
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_bn_act_pool_bwd(const float* y, const float* mean, const float* rstd, const float* bias,
                                      const float* dout, float* gz, double* sums_bwd,
                                      int B, int M, int F, int p, int pool, int act, void* stream)
{
    if (gz) return bn_act_pool_bwd_launch<0, float>(y, mean, rstd, bias, dout, gz, sums_bwd, 1.0, B, M, F, p, pool, act, STREAM);
    return bn_act_pool_bwd_launch<1, float>(y, mean, rstd, bias, dout, nullptr, sums_bwd, 1.0, B, M, F, p, pool, act, STREAM);
}

extern "C" int dsb200_bn_act_pool_bwd_bf16(const void* y, const float* mean, const float* rstd, const float* bias,
                                           const void* dout, void* gz, double* sums_bwd,
                                           int B, int M, int F, int p, int pool, int act, void* stream)
{
    if (gz) return bn_act_pool_bwd_launch<0, bf16>((const bf16*)y, mean, rstd, bias, (const bf16*)dout, (bf16*)gz, sums_bwd, 1.0, B, M, F, p, pool, act, STREAM);
    return bn_act_pool_bwd_launch<1, bf16>((const bf16*)y, mean, rstd, bias, (const bf16*)dout, nullptr, sums_bwd, 1.0, B, M, F, p, pool, act, STREAM);
}

extern "C" int dsb200_bn_act_pool_bwd_dy(const float* y, const float* mean, const float* rstd, const float* bias,
                                         const float* dout, const double* sums_bwd, double count, float* dy,
                                         int B, int M, int F, int p, int pool, int act, void* stream)
{
    DSB_REQUIRE(mean && rstd && dy && count > 0, "bn_act_pool_bwd_dy: bad arguments");
    return bn_act_pool_bwd_launch<2, float>(y, mean, rstd, bias, dout, dy, const_cast<double*>(sums_bwd), count, B, M, F, p, pool, act, STREAM);
}

extern "C" int dsb200_bn_act_pool_bwd_dy_bf16(const void* y, const float* mean, const float* rstd, const float* bias,
                                              const void* dout, const double* sums_bwd, double count, void* dy,
                                              int B, int M, int F, int p, int pool, int act, void* stream)
{
    DSB_REQUIRE(mean && rstd && dy && count > 0, "bn_act_pool_bwd_dy: bad arguments");
    return bn_act_pool_bwd_launch<2, bf16>((const bf16*)y, mean, rstd, bias, (const bf16*)dout, (bf16*)dy, const_cast<double*>(sums_bwd), count, B, M, F, p, pool, act, STREAM);
}

template <typename T>
static int bn_relu_pool_sums_launch(const T* out, const T* dout, const float* bias, double* sums_bwd,
                                    long long Rout, int F, void* stream)
{
    DSB_REQUIRE(out && dout && sums_bwd && Rout > 0 && F > 0, "bn_relu_pool_sums: bad arguments");
    DSB_REQUIRE(F % 4 == 0 && F <= 1024, "bn_relu_pool_sums: F must be a multiple of 4, at most 1024");
    DSB_REQUIRE(aligned16(out, dout), "bn_relu_pool_sums: tensors must be 16-byte aligned");
    const int Fv = F / 4, rpb = 256 / Fv;
    const int grid = grid_for((Rout + rpb - 1) / rpb * 256, 256, 8);
    bn_relu_pool_sums_kernel<T><<<grid, 256, 2 * (size_t)F * sizeof(float), STREAM>>>(out, dout, bias, sums_bwd, Rout, Fv);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_bn_relu_pool_sums(const float* out, const float* dout, const float* bias, double* sums_bwd,
                                        long long Rout, int F, void* stream)
{
    return bn_relu_pool_sums_launch<float>(out, dout, bias, sums_bwd, Rout, F, stream);
}

extern "C" int dsb200_bn_relu_pool_sums_bf16(const void* out, const void* dout, const float* bias, double* sums_bwd,
                                             long long Rout, int F, void* stream)
{
    return bn_relu_pool_sums_launch<bf16>((const bf16*)out, (const bf16*)dout, bias, sums_bwd, Rout, F, stream);
}

extern "C" int dsb200_bn_bwd_apply(const float* y, const float* mean, const float* rstd, const double* sums_bwd,
                                   double count, float* gz, long long R, int F, void* stream)
{
    DSB_REQUIRE(y && mean && rstd && sums_bwd && gz && R > 0 && F > 0 && count > 0, "bn_bwd_apply: bad arguments");
    bn_bwd_apply_kernel<<<grid_for(R * F, 256), 256, 0, STREAM>>>(y, mean, rstd, sums_bwd, count, gz, R, F);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_vertex_resize(const float* x, float* out, int B, int Min, int Mout, int F, float fill, void* stream)
{
    DSB_REQUIRE(x && out && B > 0 && Min > 0 && Mout > 0 && F > 0, "vertex_resize: bad arguments");
    vertex_resize_kernel<<<grid_for((long long)B * Mout * F, 256), 256, 0, STREAM>>>(x, out, B, Min, Mout, F, fill);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_gather_rows(const float* x, const int32_t* idx, long long n, int F, float* out, void* stream)
{
    DSB_REQUIRE(x && idx && out && n > 0 && F > 0, "gather_rows: bad arguments");
    gather_rows_kernel<<<grid_for(n * F, 256), 256, 0, STREAM>>>(x, idx, n, F, out);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_unpool_fwd(const float* x, float* out, int B, int M, int F, int p, int mode, void* stream)
{
    DSB_REQUIRE(x && out && B > 0 && M > 0 && F > 0 && p >= 1, "unpool_fwd: bad arguments");
    unpool_fwd_kernel<<<grid_for((long long)B * M * p * F, 256), 256, 0, STREAM>>>(x, out, (long long)B * M, F, p, mode);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_unpool_bwd(const float* dout, float* dx, int B, int M, int F, int p, int mode, void* stream)
{
    DSB_REQUIRE(dout && dx && B > 0 && M > 0 && F > 0 && p >= 1, "unpool_bwd: bad arguments");
    unpool_bwd_kernel<<<grid_for((long long)B * M * F, 256), 256, 0, STREAM>>>(dout, dx, (long long)B * M, F, p, mode);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_concat_f(const float* a, const float* b, float* out, long long R, int Fa, int Fb, void* stream)
{
    DSB_REQUIRE(a && b && out && R > 0 && Fa > 0 && Fb > 0, "concat_f: bad arguments");
    concat_f_kernel<<<grid_for(R * (Fa + Fb), 256), 256, 0, STREAM>>>(a, b, out, R, Fa, Fb);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_split_f(const float* dout, float* da, float* db, long long R, int Fa, int Fb, int accumulate_b,
                              void* stream)
{
    DSB_REQUIRE(dout && da && db && R > 0 && Fa > 0 && Fb > 0, "split_f: bad arguments");
    split_f_kernel<<<grid_for(R * (Fa + Fb), 256), 256, 0, STREAM>>>(dout, da, db, R, Fa, Fb, accumulate_b);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_stat_fwd(const float* x, float* out, int B, int M, int F, int kind, void* stream)
{
    DSB_REQUIRE(x && out && B > 0 && M > 0 && F > 0 && kind >= 0 && kind <= 3, "stat_fwd: bad arguments");
    const long long threads = (long long)B * F * 32;
    stat_fwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, STREAM>>>(x, out, B, M, F, kind);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_stat_bwd(const float* x, const float* dout, float* dx, int B, int M, int F, int kind, void* stream)
{
    DSB_REQUIRE(x && dout && dx && B > 0 && M > 0 && F > 0 && kind >= 0 && kind <= 3, "stat_bwd: bad arguments");
    const long long threads = (long long)B * F * 32;
    stat_bwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, STREAM>>>(x, dout, dx, B, M, F, kind);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_fc_fwd(const float* x, const float* W, const float* b, float* y, int N, int Min, int Mout,
                             int act, void* stream)
{
    DSB_REQUIRE(x && W && y && N > 0 && Min > 0 && Mout > 0, "fc_fwd: bad arguments");
    fc_fwd_kernel<<<(N * Mout + 255) / 256, 256, 0, STREAM>>>(x, W, b, y, N, Min, Mout, act);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_fc_bwd(const float* x, const float* W, const float* y, const float* dy, float* dx, float* dW,
                             float* db, int N, int Min, int Mout, int act, void* stream)
{
    DSB_REQUIRE(x && W && y && dy && dW && N > 0 && Min > 0 && Mout > 0, "fc_bwd: bad arguments");
    if (dx) { fc_bwd_dx_kernel<<<(N * Min + 255) / 256, 256, 0, STREAM>>>(W, y, dy, dx, N, Min, Mout, act); count_launch(); }
    fc_bwd_dw_kernel<<<((Min + 1) * Mout + 255) / 256, 256, 0, STREAM>>>(x, y, dy, dW, db, N, Min, Mout, act);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_hist_fwd(const float* x, const float* centers, const float* widths, float* hist,
                               int B, int M, int F, int bins, float scale, void* stream)
{
    DSB_REQUIRE(x && centers && widths && hist && B > 0 && M > 0 && F > 0, "hist_fwd: bad arguments");
    DSB_REQUIRE(bins > 0 && bins <= kHistBins, "hist_fwd: at most 32 bins");
    hist_fwd_kernel<<<dim3((F + 31) / 32, B), 256, 0, STREAM>>>(x, centers, widths, hist, M, F, bins, scale / (float)M);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_hist_bwd(const float* x, const float* centers, const float* widths, const float* dhist, float* dx,
                               float* dcenters, float* dwidths, int B, int M, int F, int bins, float scale, void* stream)
{
    DSB_REQUIRE(x && centers && widths && dhist && dx && dcenters && dwidths && B > 0 && M > 0 && F > 0, "hist_bwd: bad arguments");
    DSB_REQUIRE(bins > 0 && bins <= kHistBins, "hist_bwd: at most 32 bins");
    hist_bwd_kernel<<<dim3((F + 31) / 32, B), 256, 0, STREAM>>>(x, centers, widths, dhist, dx, dcenters, dwidths, M, F, bins,
                                                               scale / (float)M);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_softmax_ce(const float* logits, const int32_t* labels, float* dlogits, float* loss_out,
                                 long long R, int C, float grad_scale, void* stream)
{
    DSB_REQUIRE(logits && labels && loss_out && R > 0 && C > 0, "softmax_ce: bad arguments");
    softmax_ce_kernel<<<grid_for(R, 256, 4), 256, 0, STREAM>>>(logits, labels, dlogits, loss_out, R, C, grad_scale, nullptr);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_softmax_ce_weighted(const float* logits, const int32_t* labels, const float* class_weights, float* dlogits,
                                          float* loss_out, long long R, int C, float grad_scale, void* stream)
{
    DSB_REQUIRE(logits && labels && class_weights && loss_out && R > 0 && C > 0, "softmax_ce_weighted: bad arguments");
    softmax_ce_kernel<<<grid_for(R, 256, 4), 256, 0, STREAM>>>(logits, labels, dlogits, loss_out, R, C, grad_scale, class_weights);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_mse(const float* pred, const float* target, float* dpred, float* loss_out,
                          long long n, float grad_scale, void* stream)
{
    DSB_REQUIRE(pred && target && loss_out && n > 0, "mse: bad arguments");
    mse_kernel<<<grid_for(n, 256, 4), 256, 0, STREAM>>>(pred, target, dpred, loss_out, n, grad_scale);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_softmax_argmax(const float* logits, float* probs, int32_t* pred, long long R, int C, void* stream)
{
    DSB_REQUIRE(logits && R > 0 && C > 0, "softmax_argmax: bad arguments");
    softmax_argmax_kernel<<<grid_for(R, 256, 4), 256, 0, STREAM>>>(logits, probs, pred, R, C);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_l2_reg(const float* w, const float* coef, float* g, long long n, float* loss_out, void* stream)
{
    DSB_REQUIRE(w && coef && n > 0, "l2_reg: bad arguments");
    l2_reg_kernel<<<grid_for(n, 256, 2), 256, 0, STREAM>>>(w, coef, g, n, loss_out);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_optimizer_step(int kind, float* w, const float* g, float* s1, float* s2,
                                     const float* hyper, long long n, void* stream)
{
    DSB_REQUIRE(w && g && hyper && n > 0 && kind >= 0 && kind <= 3, "optimizer_step: bad arguments");
    DSB_REQUIRE(kind == DSB200_OPT_SGD || s1, "optimizer_step: missing slot 1");
    DSB_REQUIRE((kind != DSB200_OPT_ADAM && kind != DSB200_OPT_RMSPROP) || s2, "optimizer_step: missing slot 2");
    optimizer_kernel<<<grid_for(n, 256, 2), 256, 0, STREAM>>>(kind, w, g, s1, s2, hyper, n);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_fill(float* x, float v, long long n, void* stream)
{
    DSB_REQUIRE(x && n > 0, "fill: bad arguments");
    fill_kernel<<<grid_for(n, 256), 256, 0, STREAM>>>(x, v, n);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_axpy(const float* x, float* y, float a, long long n, void* stream)
{
    DSB_REQUIRE(x && y && n > 0, "axpy: bad arguments");
    axpy_kernel<<<grid_for(n, 256), 256, 0, STREAM>>>(x, y, a, n);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_mul_channel(const float* a, const float* x, int F, int ch, float* out, long long n, void* stream)
{
    DSB_REQUIRE(a && x && out && n > 0 && F > 0 && ch >= 0 && ch < F, "mul_channel: bad arguments");
    mul_channel_kernel<<<grid_for(n, 256), 256, 0, STREAM>>>(a, x, F, ch, out, n);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_scale(float* x, float a, long long n, void* stream)
{
    DSB_REQUIRE(x && n > 0, "scale: bad arguments");
    scale_kernel<<<grid_for(n, 256), 256, 0, STREAM>>>(x, a, n);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_dropout(const float* x, const float* u, float keep, float* out, long long n, void* stream)
{
    DSB_REQUIRE(x && u && out && n > 0 && keep > 0.f && keep <= 1.f, "dropout: bad arguments");
    DSB_REQUIRE(x != out, "dropout: out must not alias x");
    dropout_kernel<<<grid_for(n, 256), 256, 0, STREAM>>>(x, u, keep, out, n);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_cast_f32_to_bf16(const float* x, void* out, long long n, void* stream)
{
    DSB_REQUIRE(x && out && n > 0, "cast: bad arguments");
    DSB_REQUIRE(aligned16(x) && (((uintptr_t)out) & 7) == 0, "cast: tensors must be 16-byte aligned");
    cast_kernel<float, bf16><<<grid_for(n / 4 + 1, 256), 256, 0, STREAM>>>(x, (bf16*)out, n / 4, n);
    DSB_LAUNCH_CHECK();
}

extern "C" int dsb200_cast_bf16_to_f32(const void* x, float* out, long long n, void* stream)
{
    DSB_REQUIRE(x && out && n > 0, "cast: bad arguments");
    DSB_REQUIRE(aligned16(out) && (((uintptr_t)x) & 7) == 0, "cast: tensors must be 16-byte aligned");
    cast_kernel<bf16, float><<<grid_for(n / 4 + 1, 256), 256, 0, STREAM>>>((const bf16*)x, out, n / 4, n);
    DSB_LAUNCH_CHECK();
}
