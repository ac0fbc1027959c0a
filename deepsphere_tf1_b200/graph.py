"""Device-side packing of the rescaled Laplacians for the sparse products of the hot path.

The reference turns each scipy Laplacian into a `tf.SparseTensor` constant, reordered row-major
and cast to float32 (deepsphere/models.py:1044-1047).  Here each level becomes a `DeviceGraph`:
fixed-width ELL (HEALPix 8-neighbour / icosahedral graphs: <= 9 / 7 entries per row) or CSR
(irregular kNN graphs), together with the same packing of L^T for the backward pass (L is
symmetric only up to 1 ulp because of `D12 * W * D12`, reference utils.py:240).
"""
import numpy as np
import torch
from scipy import sparse

ELL_MAX_WIDTH = 16


class DeviceGraph(object):
    """One Laplacian level on the GPU.  `.fwd` / `.bwd` = (rowptr | None, col, val, width)."""

    def __init__(self, L, device, force_csr=False):
        L = sparse.csr_matrix(L, dtype=np.float32)
        L.sum_duplicates()
        L.sort_indices()                                   # row-major, ascending columns (tf.sparse_reorder)
        if L.shape[0] != L.shape[1]:
            raise ValueError('Laplacian must be square')
        self.M = L.shape[0]
        self.nnz = int(L.nnz)
        self.device = device
        counts = np.diff(L.indptr)
        self.max_degree = int(counts.max()) if self.M else 0
        self.is_ell = (not force_csr) and self.max_degree <= ELL_MAX_WIDTH
        self.fwd = self._pack(L)
        LT = sparse.csr_matrix(L.T)
        LT.sort_indices()
        self.symmetric_pattern = (LT.nnz == L.nnz and np.array_equal(LT.indptr, L.indptr)
                                  and np.array_equal(LT.indices, L.indices))
        if self.symmetric_pattern and np.array_equal(LT.data, L.data):
            self.bwd = self.fwd
        else:
            self.bwd = self._pack(LT)

    def _pack(self, A):
        if self.is_ell:
            width = max(self.max_degree, int(np.diff(A.indptr).max()))
            # padding slots: (col = row, val = 0) -- a harmless gather, so kernels need no early exit
            col = np.repeat(np.arange(self.M, dtype=np.int32)[:, None], width, axis=1)
            val = np.zeros((self.M, width), dtype=np.float32)
            counts = np.diff(A.indptr)
            rows = np.repeat(np.arange(self.M), counts)
            slot = np.arange(A.nnz) - np.repeat(A.indptr[:-1], counts)
            col[rows, slot] = A.indices
            val[rows, slot] = A.data
            return (None, torch.from_numpy(col).to(self.device), torch.from_numpy(val).to(self.device), width)
        return (torch.from_numpy(A.indptr.astype(np.int32)).to(self.device),
                torch.from_numpy(A.indices.astype(np.int32)).to(self.device),
                torch.from_numpy(A.data.astype(np.float32)).to(self.device), 0)

    def bytes_stored(self):
        """Bytes one sparse product reads for the matrix (8 per stored entry)."""
        return 8 * self.nnz + (0 if self.is_ell else 4 * (self.M + 1))
