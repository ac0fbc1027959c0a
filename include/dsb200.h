/*
 * dsb200.h -- C ABI of the B200-native DeepSphere ChebConv hot path (libdsb200.so).
 *
 * The reference (deepsphere/deepsphere-tf1) is pure Python on TensorFlow 1.10 and has no
 * FFI; its only operator seam is the `getattr` binding in cgcnn.__init__
 * (deepsphere/models.py:1017-1020: self.filter / self.pool / self.unpool / self.activation)
 * plus the TF ops those methods call.  Each entry point below replaces one of those TF op
 * call sites (cited per function); a maintainer binds them with ctypes.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer; the caller owns all buffers (the library never
 *    allocates or frees device memory on the data path);
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no host sync,
 *    no global state => re-entrant across streams and CUDA-graph capturable;
 *  - return value: 0 on success, otherwise a cudaError_t (or -1 for invalid arguments);
 *    dsb200_last_error() returns a static message for the last failure on this thread;
 *  - activations are fp32 in the reference's own layout  x[b][m][f]  ("BMF": sample b, vertex
 *    m, feature f; models.py:675) -- a row-major matrix [R, F] with R = B*M rows for the dense
 *    contraction, and B independent row-major [M, F] matrices for the sparse products;
 *  - Chebyshev weights keep the reference layout W[Fin*K, Fout], row = fin*K + k
 *    (models.py:1062-1066);
 *  - the rescaled Laplacian is passed as ELL (col[M*width], val[M*width], row-major, unused
 *    slots at the end of a row hold col = the row itself, val = 0) or as CSR (rowptr[M+1],
 *    col[nnz], val[nnz]).
 */
#ifndef DSB200_H
#define DSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSB200_ABI_VERSION 2

/* activation ids (models.py:1020 `getattr(tf.nn, activation)`) */
enum { DSB200_ACT_IDENTITY = 0, DSB200_ACT_RELU = 1, DSB200_ACT_ELU = 2, DSB200_ACT_TANH = 3,
       DSB200_ACT_SIGMOID = 4, DSB200_ACT_LEAKY_RELU = 5, DSB200_ACT_SOFTPLUS = 6, DSB200_ACT_RELU6 = 7 };
/* pooling ids (models.py:1018 `getattr(self, 'pool_' + pool)`) */
enum { DSB200_POOL_MAX = 0, DSB200_POOL_AVERAGE = 1 };
/* statistics ids (models.py:1249-1270) */
enum { DSB200_STAT_MEAN = 0, DSB200_STAT_VAR = 1, DSB200_STAT_MEANVAR = 2, DSB200_STAT_MAX = 3 };
/* optimizer ids (tf.train.*Optimizer passed through params['optimizer']) */
enum { DSB200_OPT_SGD = 0, DSB200_OPT_MOMENTUM = 1, DSB200_OPT_ADAM = 2, DSB200_OPT_RMSPROP = 3 };
/* unpooling modes (models.py:1347-1375) */
enum { DSB200_UNPOOL_REPEAT = 0, DSB200_UNPOOL_ZEROFILL = 1 };

int dsb200_abi_version(void);
const char* dsb200_last_error(void);
/* number of CUDA kernels this library has launched in this process (all threads) */
long long dsb200_kernel_launches(void);

/* ---- K1: one sparse-Laplacian product with fused recurrence arithmetic, for all B samples
 *      out[b] = alpha * L x[b] + beta * y[b] + gamma * z[b]    (y, z NULL when beta/gamma == 0;
 *      out may alias y or z)
 * replaces tf.sparse_tensor_dense_matmul + `2*..-x0` (models.py:1056,1059); with
 * (alpha,beta)=(1,0) it is X1 = L X0, with (2,-1) it is X_k = 2 L X_{k-1} - X_{k-2}; the
 * backward pass uses it with L^T for the adjoint recurrence (autodiff of 1056-1061). */
int dsb200_spmm_ell(const int32_t* col, const float* val, int width, int B, int M, int F,
                    const float* x, const float* y, const float* z, float* out,
                    float alpha, float beta, float gamma, void* stream);
int dsb200_spmm_csr(const int32_t* rowptr, const int32_t* col, const float* val, int B, int M, int F,
                    const float* x, const float* y, const float* z, float* out,
                    float alpha, float beta, float gamma, void* stream);

/* ---- all K Chebyshev orders of one layer (models.py:1049-1061).  x0 [B,M,F]; basis receives
 * orders 1..K-1 as [K-1][B][M][F] (order 0 is x0 itself, never copied).  ELL if rowptr==NULL. */
int dsb200_cheb_basis(const int32_t* rowptr, const int32_t* col, const float* val, int width,
                      int B, int M, int F, int K, const float* x0, float* basis, void* stream);
/* adjoint of the recurrence: G [K][B][M][F] holds dLoss/dX_k on entry; on exit G[0] holds
 * dLoss/dX_0.  (colT, valT / rowptrT) describe L^T.  Overwrites G[1..K-1]. */
int dsb200_cheb_basis_bwd(const int32_t* rowptrT, const int32_t* colT, const float* valT, int width,
                          int B, int M, int F, int K, float* G, void* stream);

/* ---- K1 with PACKED ELL records (the default path of the host code).  rec int2[width][M]: rec[j*M + m] = (col, bit
 * pattern of val) of slot j of row m -- slot-major, so the rows a warp works on read ONE contiguous piece per slot
 * (one L1 wavefront instead of ~6 with the row-major col[] / val[] arrays).  dsb200_ell_pack builds it on the device
 * from the row-major arrays; the *_ellp entry points take both (the row-major arrays serve the shapes the record
 * kernel does not cover: F % 4 != 0, unaligned pointers, the one-launch recurrence of the coarsest levels).
 * Same operators and call sites as dsb200_spmm_ell / dsb200_cheb_basis / dsb200_cheb_basis_bwd (models.py:1049-1061). */
int dsb200_ell_pack(const int32_t* col, const float* val, int width, int M, void* rec, void* stream);
int dsb200_spmm_ellp(const void* rec, const int32_t* col, const float* val, int width, int B, int M, int F,
                     const float* x, const float* y, const float* z, float* out,
                     float alpha, float beta, float gamma, void* stream);
int dsb200_cheb_basis_ellp(const void* rec, const int32_t* col, const float* val, int width,
                           int B, int M, int F, int K, const float* x0, float* basis, void* stream);
int dsb200_cheb_basis_bwd_ellp(const void* recT, const int32_t* colT, const float* valT, int width,
                               int B, int M, int F, int K, float* G, void* stream);

/* ---- K1, bulk-asynchronous pipelined (csrc/sparse_pipe.cu; the default for F in {16, 32, 64}, ELL width 7 / 9, when the
 * graph has a tile plan).  Same operator as dsb200_spmm_ell (models.py:1056-1061).  A work item = (row tile of TR
 * consecutive rows, sample): the x / y / z rows of the tile arrive by cp.async.bulk (TMA) on an mbarrier, the halo rows by
 * 16-byte cp.async, several items in flight per CTA; the arithmetic reads shared memory only.
 * tile_recs int2[ntiles][width][TR] = (LOCAL column, value bits): local column < TR = row of the tile, else TR + halo slot;
 * tile_ext int32[ntiles][HC] = global row of every halo slot; tile_nh int32[ntiles] = halo rows per tile
 * (graph.py: PipePlan).  `dense` carries two flags.  Bit 1: walk the items from the end of memory to its start (same items, same
 * results; a launch that starts where its producer stopped finds the producer's last rows in L2).  Bit 0: DENSE-GROUP records instead -- per tile [18][TR/4] 16-byte pieces: a group = 4
 * consecutive rows and the <= 16 distinct local columns they touch; pieces 0..15 = the group's dense 4 x 16 weight block
 * (row r ^ (g & 1) of group g, columns 4q..4q+3 at piece 4r + q), pieces 16, 17 = the 16 uint16 local columns.  Each distinct neighbour is read
 * once per group (16 reads for 4 rows instead of 36); sums run in column order, so results equal the record form to
 * rounding (~1e-7), not bit for bit.  out must not alias x (it may alias y or z); pointers 16-byte aligned. */
int dsb200_spmm_pipe(const void* tile_recs, const int32_t* tile_ext, const int32_t* tile_nh,
                     int ntiles, int TR, int HC, int width, int dense, int B, int M, int F,
                     const float* x, const float* y, const float* z, float* out,
                     float alpha, float beta, float gamma, void* stream);

/* ---- K1 fused: ALL K - 1 sparse products of a layer in one launch (csrc/cheb_tile.cu; HEALPix NESTED graphs whose vertex
 * list is whole 256-pixel tiles: the cosmo patches).  Replaces the K - 1 tf.sparse_tensor_dense_matmul calls of
 * cgcnn.chebyshev5 (models.py:1056-1061) -- or of cgcnn.monomials (models.py:1099-1102) when monomial != 0 -- and, with
 * backward != 0, the TF autodiff of those lines.  S = K - 1 in 1..7, T = 16, M = ntiles * 256, F % 16 == 0.
 *   tile_p2v int32[ntiles][P][P], P = T + 2 S: vertex shown at every position of the tile's flat image (-1 = none);
 *   tile_w  float[ntiles][P-2][P-2][12]: the 3 x 3 weights of L (backward: of L^T) around every computed position, entry
 *           3 (dy + 1) + (dx + 1) multiplies the value at (row + dy, column + dx); entries 9..11 are padding
 *   (graph.py: StencilPlan builds and verifies both from the ELL arrays).
 * forward : x = X_0 [B][M][F],            out = orders 1..K-1 [K-1][B][M][F]   (same result as dsb200_cheb_basis)
 * backward: x = G [K][B][M][F] (read only), out = dLoss/dX_0 [B][M][F]          (same result as dsb200_cheb_basis_bwd) */
int dsb200_cheb_tile(const int32_t* tile_p2v, const float* tile_w, int ntiles, int T, int S, int B, int M, int F,
                     int backward, int monomial, const float* x, float* out, void* stream);

/* ---- K2: dense contraction over (fin, k)  (tf.matmul, models.py:1062-1078)
 *   Y[r, fo] = sum_{k, fi} X_k[r, fi] * W[fi*K + k, fo],   r < R = B*M
 * x0 = order 0 [R,Fin], xk = orders 1..K-1 [K-1][R][Fin] (may be NULL when K == 1).
 * If bn_sums != NULL (double[2*Fout], zeroed by the caller) the epilogue also accumulates
 * sum_r Y and sum_r Y^2 per output feature (tf.nn.moments inside
 * tf.layers.batch_normalization, models.py:1197-1203). */
int dsb200_cheb_contract_fwd(const float* x0, const float* xk, const float* W, float* Y,
                             long long R, int Fin, int K, int Fout, double* bn_sums, void* workspace, void* stream);
/* dX_k[r, fi] = sum_fo dY[r, fo] * W[fi*K + k, fo]  -> G [K][R][Fin]  (autodiff of 1077) */
int dsb200_cheb_contract_bwd_data(const float* dY, const float* W, float* G,
                                  long long R, int Fin, int K, int Fout, void* workspace, void* stream);
/* The same contraction with the HEALPix max-pool over 4 consecutive vertices (models.py:1139-1142 with p = 4) fused into the
 * epilogue: Ypool [R/4, Fout] = max over every 4 consecutive rows of Y.  tf.layers.batch_normalization without scale /
 * shift (models.py:1197), the bias and every activation of the library are non-decreasing per feature, so
 * pool(act(bn(y) + b)) = act(bn(pool(y)) + b): the normalise / activate pass (dsb200_bn_act_pool_fwd with p = 1) then reads a
 * quarter of the rows.  Y may be NULL (inference: the un-pooled tensor is never written); bn_sums as above, over ALL rows.
 * Tensor-core kernels only: dsb200_contract_pool4_supported(R, Fin, K, Fout) != 0. */
int dsb200_cheb_contract_fwd_pool4(const float* x0, const float* xk, const float* W, float* Y, float* Ypool,
                                   long long R, int Fin, int K, int Fout, double* bn_sums, void* workspace, void* stream);
int dsb200_contract_pool4_supported(long long R, int Fin, int K, int Fout);
/* `workspace` (device, 16-byte aligned, dsb200_contract_workspace_bytes(Fin, K, Fout) bytes, may be NULL):
 * scratch for the TF32 hi / lo split weight image of the tensor-core path; without it every CTA splits
 * the weights itself (slower for small layers). */
long long dsb200_contract_workspace_bytes(int Fin, int K, int Fout);
/* dW[fi*K + k, fo] += sum_r X_k[r, fi] * dY[r, fo]  (dW must be zeroed by the caller) */
int dsb200_cheb_contract_bwd_weight(const float* x0, const float* xk, const float* dY, float* dW,
                                    long long R, int Fin, int K, int Fout, void* stream);

/* The three contraction entry points run on the tcgen05 tensor cores (3xTF32 error-compensated,
 * FP32 accumulate in TMEM) when the shape allows (Fin, Fout multiples of 8, R >= 128) and fall
 * back to FP32 FFMA kernels otherwise; dsb200_set_tensor_cores(0) forces the FFMA kernels.
 * Returns the previous setting. */
int dsb200_set_tensor_cores(int on);

/* ---- ChebConv layers with a tiny reduction depth K*Fin <= 8 (first layer of the cosmo net: Fin = 1):
 * the filter output y is Fout/(K*Fin) times larger than the basis, so it is never materialised.
 * Replaces, for such layers, the whole of models.py:1062-1078 + 1194-1203 + 1122-1163 and their
 * gradients.  dsb200_smallconv_supported says whether a shape qualifies. */
int dsb200_smallconv_supported(int Fin, int K, int Fout);
/* gram (double[Q + Q*Q], Q = K*Fin, zeroed by the caller): column sums and Gram matrix of the basis
 * over all R rows; index q = k*Fin + fi.  Summed over ranks by the caller for data parallelism. */
int dsb200_cheb_gram(const float* x0, const float* xk, long long R, int Fin, int K, double* gram, void* stream);
/* batch-norm sums (double[2*Fout]: sum y, sum y^2) of y = Xcat W from the Gram matrix */
int dsb200_bn_sums_from_gram(const double* gram, const float* W, int Fin, int K, int Fout, double* sums, void* stream);
/* out [B, M/p, Fout] = pool(act((Xcat W - mean) * rstd + bias)); mean/rstd/bias may be NULL.
 * interleaved != 0: the basis planes are in the sample-interleaved layout [M, B, Fin] (the input was
 * passed through dsb200_interleave before the recurrence, which then ran on 1 "sample" of B*Fin features);
 * outputs / dout keep the classic [B, M/p, Fout] layout. */
int dsb200_smallconv_fwd(const float* x0, const float* xk, const float* W, const float* mean, const float* rstd,
                         const float* bias, float* out, int B, int M, int Fin, int K, int Fout, int p, int pool,
                         int act, int interleaved, uint8_t* argmax, void* stream);
/* relu + max pooling over 4 vertices (dsb200_smallconv_has_argmax != 0): the forward can record the winning
 * sibling (argmax, uint8 [B, M/p, Fout], may be NULL); given `out` and `argmax` the backward needs no
 * recomputation of y (gz = dout * (out > 0), xhat = out - bias). */
int dsb200_smallconv_has_argmax(int Fout, int p, int pool, int act);
/* out [M, B, F] = transpose of x [B, M, F] over (sample, vertex) */
int dsb200_interleave(const float* x, float* out, int B, int M, int F, void* stream);
/* acc (double[(Q+2)*Fout], zeroed): A[q][f] = sum_r X_q[r] gz[r,f], then S1[f] = sum gz, S2[f] = sum gz*xhat,
 * with gz = dLoss/d(normalised + bias) recomputed from the basis and dout [B, M/p, Fout] */
int dsb200_smallconv_bwd(const float* x0, const float* xk, const float* W, const float* mean, const float* rstd,
                         const float* bias, const float* dout, double* acc, int B, int M, int Fin, int K, int Fout,
                         int p, int pool, int act, int interleaved, const float* out, const uint8_t* argmax,
                         void* stream);
/* dW (+=) and dbias (=, may be NULL) from acc (already summed over ranks) and the Gram matrix; with
 * mean == NULL there is no batch norm and dW += A. */
int dsb200_smallconv_bwd_finalize(const double* acc, const double* gram, const float* W, const float* mean,
                                  const float* rstd, double count, int Fin, int K, int Fout, float* dW,
                                  float* dbias, void* stream);

/* ---- batch-norm statistics -> mean / rstd, moving-average update (models.py:1194-1203,
 * UPDATE_OPS at 840-842).  sums = double[2*F] (sum, sum of squares over `count` rows, already
 * all-reduced over ranks).  moving_* may be NULL.  momentum 0.9, eps 1e-5 in the reference. */
int dsb200_bn_finalize(const double* sums, double count, int F, float eps, float momentum,
                       float* mean, float* rstd, float* moving_mean, float* moving_var, void* stream);
/* evaluation mode: mean/rstd from the moving statistics */
int dsb200_bn_from_moving(const float* moving_mean, const float* moving_var, int F, float eps,
                          float* mean, float* rstd, void* stream);

/* ---- K3: fused  normalise -> + bias -> activation -> pool  (models.py:1235-1243)
 *   y [B,M,F] -> out [B,M/p,F];   z = (y - mean)*rstd + bias (mean/rstd NULL = no batch norm,
 *   bias NULL = none); pooling over p consecutive vertices (HEALPix nested order). */
int dsb200_bn_act_pool_fwd(const float* y, const float* mean, const float* rstd, const float* bias,
                           float* out, int B, int M, int F, int p, int pool, int act, void* stream);
/* backward of the same chain: gz [B,M,F] = dLoss/dz, and sums_bwd (double[2*F], zeroed by the
 * caller) += (sum gz, sum gz * xhat) per feature; sum gz is also the bias gradient.  gz == NULL:
 * accumulate the sums only (pass 1 of the batch-norm backward; nothing is written). */
int dsb200_bn_act_pool_bwd(const float* y, const float* mean, const float* rstd, const float* bias,
                           const float* dout, float* gz, double* sums_bwd,
                           int B, int M, int F, int p, int pool, int act, void* stream);
/* pass 2 of the batch-norm backward without a stored gz: dy = rstd * (gz - S1/count - xhat * S2/count)
 * with gz recomputed from y and dout; sums_bwd already summed over ranks */
int dsb200_bn_act_pool_bwd_dy(const float* y, const float* mean, const float* rstd, const float* bias,
                              const float* dout, const double* sums_bwd, double count, float* dy,
                              int B, int M, int F, int p, int pool, int act, void* stream);
/* pass 1 of the batch-norm backward for relu + max pool (or relu without pooling), from the POOLED
 * tensors only: out [Rout, F] (the layer output saved by the forward pass) and dout [Rout, F];
 * sums_bwd (double[2*F], zeroed by the caller) += (sum gz, sum gz * xhat), using
 * gz != 0 only where out > 0, and xhat = out - bias there (models.py:1235-1243 autodiff). */
int dsb200_bn_relu_pool_sums(const float* out, const float* dout, const float* bias, double* sums_bwd,
                             long long Rout, int F, void* stream);
/* dy = rstd * (gz - S1/count - xhat * S2/count), in place on gz (batch-norm backward);
 * dbias (float[F], may be NULL) receives S1. */
int dsb200_bn_bwd_apply(const float* y, const float* mean, const float* rstd, const double* sums_bwd,
                        double count, float* gz, long long R, int F, void* stream);
/* dbias[f] = (float) sums_bwd[f] */
int dsb200_sums_to_float(const double* sums, float* out, int n, void* stream);

/* ---- vertex-count change: out[b, m, f] = m < Min ? x[b, m, f] : fill,  m < Mout.
 * Mout < Min is the icosahedral "pool" x[:, :p, :] (models.py:1138,1157); Mout > Min with
 * fill = 1 is the icosahedral unpool tf.pad(..., constant_values=1) (models.py:1356-1358);
 * fill = 0 gives the backward of the slice. */
int dsb200_vertex_resize(const float* x, float* out, int B, int Min, int Mout, int F, float fill, void* stream);
/* out[i, :] = x[idx[i], :] for i < n: packs the boundary rows a neighbouring rank needs into the send buffer of
 * the halo exchange of the vertex-sharded mode (new; the reference is single-device). */
int dsb200_gather_rows(const float* x, const int32_t* idx, long long n, int F, float* out, void* stream);
/* ---- HEALPix unpooling (models.py:1362-1364 repeat, 1372-1375 value then p-1 zeros) */
int dsb200_unpool_fwd(const float* x, float* out, int B, int M, int F, int p, int mode, void* stream);
int dsb200_unpool_bwd(const float* dout, float* dx, int B, int M, int F, int p, int mode, void* stream);

/* ---- feature concat of two tensors along F (tf.concat axis=-1... models.py:1311) */
int dsb200_concat_f(const float* a, const float* b, float* out, long long R, int Fa, int Fb, void* stream);
/* da = dout[:, :Fa];  db (+)= dout[:, Fa:]  (accumulate_b != 0 adds into db) */
int dsb200_split_f(const float* dout, float* da, float* db, long long R, int Fa, int Fb, int accumulate_b,
                   void* stream);

/* ---- K4: statistical head over vertices (models.py:1249-1270): x [B,M,F] -> out [B, F or 2F] */
int dsb200_stat_fwd(const float* x, float* out, int B, int M, int F, int kind, void* stream);
int dsb200_stat_bwd(const float* x, const float* dout, float* dx, int B, int M, int F, int kind, void* stream);

/* ---- small dense layers of the head (cgcnn.fc, models.py:1205-1217): y = act(x W + b) */
int dsb200_fc_fwd(const float* x, const float* W, const float* b, float* y, int N, int Min, int Mout,
                  int act, void* stream);
/* dz = dy * act'(z) is computed from y (= act(z)) for relu / identity only; dx = dz W^T,
 * dW += x^T dz, db += sum dz.  dx may be NULL. */
int dsb200_fc_bwd(const float* x, const float* W, const float* y, const float* dy, float* dx, float* dW,
                  float* db, int N, int Min, int Mout, int act, void* stream);

/* ---- losses (models.py:777-805).  logits [R, C] row-major, labels int32 [R]; loss_out[0] +=
 * mean CE, dlogits = (softmax - onehot) * grad_scale / R (dlogits may be NULL).
 * MSE: pred/target [n]. */
int dsb200_softmax_ce(const float* logits, const int32_t* labels, float* dlogits, float* loss_out,
                      long long R, int C, float grad_scale, void* stream);
/* class-weighted variant (`weighted=True`, models.py:795-802): row r counts class_weights[labels[r]] times, the mean is
 * still over R rows */
int dsb200_softmax_ce_weighted(const float* logits, const int32_t* labels, const float* class_weights, float* dlogits,
                               float* loss_out, long long R, int C, float grad_scale, void* stream);
int dsb200_mse(const float* pred, const float* target, float* dpred, float* loss_out,
               long long n, float grad_scale, void* stream);
/* softmax probabilities / argmax (models.py:750-763); probs or pred may be NULL */
int dsb200_softmax_argmax(const float* logits, float* probs, int32_t* pred, long long R, int C, void* stream);

/* ---- learned histogram head (cgcnn.learned_histogram, models.py:1166-1192; statistics='histogram', models.py:1263-1266)
 *   hist[b, f, j] = scale * mean_m relu(1 - |x[b,m,f] - centers[f,j]| * |widths[f,j]|),  scale = bins / initial_range / 4
 * x [B,M,F]; centers / widths [F, bins]; hist [B, F*bins] (feature-major, as the reference reshapes it); bins <= 32.
 * Backward: dx [B,M,F] is written, dcenters / dwidths [F, bins] are ACCUMULATED (caller zeroes). */
int dsb200_hist_fwd(const float* x, const float* centers, const float* widths, float* hist,
                    int B, int M, int F, int bins, float scale, void* stream);
int dsb200_hist_bwd(const float* x, const float* centers, const float* widths, const float* dhist, float* dx,
                    float* dcenters, float* dwidths, int B, int M, int F, int bins, float scale, void* stream);

/* ---- regularisation term  sum_i ||W_i||^2 / 2 / std_i^2  over a flat parameter buffer with a
 * per-element coefficient (models.py:806-808, 866): loss_out[0] += sum coef*w^2/2 and
 * g += coef * w */
int dsb200_l2_reg(const float* w, const float* coef, float* g, long long n, float* loss_out, void* stream);

/* ---- optimizer step over a flat parameter buffer (tf.train.*Optimizer.apply_gradients,
 * models.py:831-832).  hyper (DEVICE, float[8]): [0] lr, [1] beta1 | momentum | decay,
 * [2] beta2 | rmsprop momentum, [3] epsilon, [4] lr_t (Adam bias-corrected step size),
 * [5] gradient scale (1/world for data parallel sums), [6..7] unused.
 * s1/s2 = optimizer slots (Adam m, v; Momentum accumulator; RMSProp ms, mom). */
int dsb200_optimizer_step(int kind, float* w, const float* g, float* s1, float* s2,
                          const float* hyper, long long n, void* stream);

/* ---- misc */
int dsb200_fill(float* x, float v, long long n, void* stream);
int dsb200_axpy(const float* x, float* y, float a, long long n, void* stream); /* y += a*x */
int dsb200_scale(float* x, float a, long long n, void* stream);               /* x *= a */
/* out[i] = a[i] * x[i*F + ch], i < n: a feature channel of the input batch as a mask (masked MSE, models.py:785-787) */
int dsb200_mul_channel(const float* a, const float* x, int F, int ch, float* out, long long n, void* stream);
/* tf.nn.dropout on the fc layers (models.py:1279-1280): out = u < keep ? x / keep : 0 with u a
 * caller-supplied uniform [0,1) tensor; the same call applied to dout is the backward. */
int dsb200_dropout(const float* x, const float* u, float keep, float* out, long long n, void* stream);

/* ---- halo exchange of the vertex-sharded mode over NVLink peer memory (csrc/peer_halo.cu; new).  Every rank owns a mailbox
 * (dsb200_peer_halo_create, `capacity` floats per half) that its peers open with dsb200_peer_comm_open.  One exchange = one
 * kernel: push the boundary rows into the neighbours' mailboxes, publish a sequence number, wait for the neighbours', copy the
 * mailbox into the ghost rows rows[Ml .. Ml + H).  meta int32 (device): [0] sends, [1] receives, per send (peer, rows, first ghost
 * row at the peer, first entry in idx), per receive (peer); idx int32 (device): owned rows to send, per neighbour ascending.
 * HOST pointers: comm_out, handle, comms.  16-byte pieces when F % 4 == 0 and rows is 16-byte aligned, single floats otherwise. */
int dsb200_peer_halo_create(long long capacity, void** comm_out, void* handle_out64);
int dsb200_peer_halo_exchange(void* const* comms, int rank, int world, float* rows, int F, int Ml, int H,
                              const int32_t* meta, const int32_t* idx, int max_rows, void* stream);

/* ---- device-side HEALPix graph / Laplacian construction (csrc/healpix_build.cu; replaces the host path of
 * deepsphere/utils.py:25-133, 231-242, 379-385 and healpy's pix2vec / get_all_neighbours, utils.py:56, 63).  The kernels follow
 * the numpy path operation by operation (explicitly rounded float32 / float64 arithmetic), the host does the three array
 * operations in between whose rounding belongs to numpy's libm (graph_build.py): mean -> kernel width, exp -> weights, power ->
 * d^-1/2.  ids (int64[M]) = NESTED pixel of every vertex or NULL (vertex v = pixel v); inverse (int32[12 nside^2]) = vertex of
 * every pixel (-1 = none) or NULL with ids == NULL.
 *   hpx_coords      coords float[M][3] = float32(pix2vec(nside, pixel, nest=True))
 *   hpx_neighbours  nb int32[M][8] = vertex of the SW, W, NW, N, NE, E, SE, S neighbour (-1 = none), d2 float[M][8] = squared
 *                   chord distance ((dx^2 + dy^2) + dz^2 in float32 (0 where nb < 0)
 *   hpx_sort_degree w float[M][8] (weights, host-computed) -> per row the entries in ascending column order (scol, sw; -1 / 0
 *                   beyond cnt[v]) and deg[v] = their sequential float32 sum (scipy: canonical CSR, W.sum(1))
 *   hpx_laplacian   ELL rows (col int32[M][width], val float[M][width], padding = (row, 0)) of I - (d12_i w_ij) d12_j
 *                   (normalized != 0; d12 = deg^-1/2 from the host) or D - W (normalized == 0), diagonal at its sorted place
 *   ell_rescale     val <- val * recip, diagonal - 1; recip = float32(1.0 / (lmax / 2)) -- scipy evaluates the `L /= lmax / 2`
 *                   of rescale_L (utils.py:379-385) as a multiplication by the reciprocal
 *   ell_transpose   valT = values of the transposed matrix on the same (symmetric) pattern; *missing counts entries whose
 *                   transposed partner is absent */
int dsb200_hpx_coords(int nside, const long long* ids, long long M, float* coords, void* stream);
int dsb200_hpx_neighbours(int nside, const long long* ids, const int32_t* inverse, long long M, const float* coords,
                          int32_t* nb, float* d2, void* stream);
int dsb200_hpx_sort_degree(const int32_t* nb, const float* w, long long M, int32_t* scol, float* sw, int32_t* cnt,
                           float* deg, void* stream);
int dsb200_hpx_laplacian(const int32_t* scol, const float* sw, const int32_t* cnt, const float* d12, const float* deg,
                         long long M, int width, int normalized, int32_t* col, float* val, void* stream);
int dsb200_ell_rescale(const int32_t* col, float* val, long long M, int width, float recip, void* stream);
int dsb200_ell_transpose(const int32_t* col, const float* val, long long M, int width, float* valT, int32_t* missing,
                         void* stream);

/* ---- device-side input pipeline (csrc/input.cu; replaces the host batch assembly of deepsphere/data.py:39-147 + feed_dict,
 * models.py:535).  out[b, j] = X[idx[b], j] + level * sample_scale[b] * N(0, 1), j < row_len: gather of the batch's samples from
 * the HBM-resident dataset plus Gaussian noise (LabeledDatasetWithNoise / GaussianNoise, data.py:95-166) from a counter-based
 * generator (Philox4x32-10, Box-Muller): the number at (b, j) depends only on (seed, offset, b, j).  sample_scale may be NULL
 * (all ones; the `all_level` option of GaussianNoise draws it per sample).  level == 0: a plain gather. */
int dsb200_gather_noise(const float* X, const int32_t* idx, int nb, long long row_len, float level,
                        const float* sample_scale, long long seed, long long offset, float* out, void* stream);
/* out[b, j] = labels[idx[b], j] (int32; row_len = 1 for one label per sample, M for dense labels) */
int dsb200_gather_labels(const int32_t* labels, const int32_t* idx, int nb, long long row_len, int32_t* out, void* stream);

/* ---- one-shot all-reduce of small double vectors over NVLink peer memory (csrc/peer_allreduce.cu; new -- replaces the
 * NCCL launches of the SyncBN sums inside the data-parallel / vertex-sharded step).  Every rank creates one
 * communication buffer and exports its IPC handle (64 bytes); the handles are exchanged by the host (torch.distributed
 * all_gather) and opened on every peer.  HOST pointers: comm_out, handle, comms (array of `world` device pointers as
 * mapped in this process, own buffer at index `rank`). */
long long dsb200_peer_comm_bytes(void);
int dsb200_peer_comm_create(void** comm_out, void* handle_out64);
int dsb200_peer_comm_open(const void* handle64, void** comm_out);
int dsb200_peer_allreduce_f64(void* const* comms, int rank, int world, double* data, int n, void* stream);

/* ---- BF16 mode (north star: "FP32 ... (1e-2 in BF16)"; cgcnn(..., dtype='bfloat16')).  Activations -- layer inputs and
 * outputs, the Chebyshev planes, Y, and every activation gradient -- are STORED as bfloat16 (void* = __nv_bfloat16*); all
 * arithmetic and accumulation is fp32, one rounding (to nearest even) per stored element.  Weights, batch-norm statistics
 * and weight gradients stay fp32.  Same operators and reference call sites as the fp32 entry points of the same name. */
/* K1 (models.py:1056-1061), csrc/sparse_pipe.cu; the tile plan is the one for rows of 2 F bytes (TR = 8192 / (2 F)) */
int dsb200_spmm_pipe_bf16(const void* tile_recs, const int32_t* tile_ext, const int32_t* tile_nh,
                          int ntiles, int TR, int HC, int width, int dense, int B, int M, int F,
                          const void* x, const void* y, const void* z, void* out,
                          float alpha, float beta, float gamma, void* stream);
/* K2 (models.py:1062-1078 and its autodiff), csrc/contract_bf16.cu: tcgen05.mma kind::f16 (BF16 x BF16 -> FP32 in TMEM) fed by
 * TMA tensor copies, no operand split.  Shapes: dsb200_contract_bf16_supported(dir = 0 forward / 1 data gradient / 2 weight
 * gradient, ...) != 0; other shapes: cast to fp32 and use the fp32 entry points.  workspace: device scratch of
 * dsb200_contract_workspace_bytes_bf16(R, Fin, K, Fout) bytes (the packed bf16 weight image), required.
 * bn_sums: statistics of the STORED (rounded) Y. */
int dsb200_contract_bf16_supported(int dir, long long R, int Fin, int K, int Fout);
long long dsb200_contract_workspace_bytes_bf16(long long R, int Fin, int K, int Fout);
int dsb200_cheb_contract_fwd_bf16(const void* x0, const void* xk, const float* W, void* Y,
                                  long long R, int Fin, int K, int Fout, double* bn_sums, void* workspace, void* stream);
/* forward with the fused 4-vertex max-pool (see dsb200_cheb_contract_fwd_pool4): dsb200_contract_bf16_supported(3, ...) != 0 */
int dsb200_cheb_contract_fwd_pool4_bf16(const void* x0, const void* xk, const float* W, void* Y, void* Ypool,
                                        long long R, int Fin, int K, int Fout, double* bn_sums, void* workspace, void* stream);
int dsb200_cheb_contract_bwd_data_bf16(const void* dY, const float* W, void* G,
                                       long long R, int Fin, int K, int Fout, void* workspace, void* stream);
int dsb200_cheb_contract_bwd_weight_bf16(const void* x0, const void* xk, const void* dY, float* dW,
                                         long long R, int Fin, int K, int Fout, void* stream);
/* K3 / K4 (models.py:1194-1203, 1122-1163, 1235-1243): the batch-norm / bias / activation / pooling chain on bf16 tensors */
int dsb200_bn_act_pool_fwd_bf16(const void* y, const float* mean, const float* rstd, const float* bias,
                                void* out, int B, int M, int F, int p, int pool, int act, void* stream);
int dsb200_bn_act_pool_bwd_bf16(const void* y, const float* mean, const float* rstd, const float* bias,
                                const void* dout, void* gz, double* sums_bwd,
                                int B, int M, int F, int p, int pool, int act, void* stream);
int dsb200_bn_act_pool_bwd_dy_bf16(const void* y, const float* mean, const float* rstd, const float* bias,
                                   const void* dout, const double* sums_bwd, double count, void* dy,
                                   int B, int M, int F, int p, int pool, int act, void* stream);
int dsb200_bn_relu_pool_sums_bf16(const void* out, const void* dout, const float* bias, double* sums_bwd,
                                  long long Rout, int F, void* stream);
/* storage conversion (operators without a bf16 kernel run in fp32 between two of these) */
int dsb200_cast_f32_to_bf16(const float* x, void* out, long long n, void* stream);
int dsb200_cast_bf16_to_f32(const void* x, float* out, long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DSB200_H */
