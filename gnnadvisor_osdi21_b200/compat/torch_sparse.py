"""Stand-in for `torch_sparse.spmm`, the CPU reference of the reference's verification mode
(GNNAdvisor/unitest.py:6,39-40).  torch_sparse is not installable offline.
spmm(index[2,E] int64, value[E], m, n, dense[n,k]) -> [m,k]; duplicate entries add up (COO semantics)."""
import torch


def spmm(index, value, m, n, matrix):
    a = torch.sparse_coo_tensor(index, value.to(matrix.dtype), (m, n))
    return torch.sparse.mm(a, matrix)
