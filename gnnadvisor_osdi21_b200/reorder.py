"""Vertex reordering -- the `rabbit` module of the reference on top of libgnna_b200.so.

reference: rabbit_module/src/reorder.cpp:235-295 (`rabbit.reorder(IntTensor[2,E]) -> IntTensor[2,E]`,
called from GNNAdvisor/dataset.py:153).  Same contract: the returned edge list has both endpoints
mapped old id -> new id and keeps the edge order of the input.
"""
import ctypes

import torch

from . import _lib


def permutation(edge_index, num_nodes=None, window=0):
    """old id -> new id (int32 CPU tensor [num_nodes]) by Rabbit Order community renumbering, on all host threads and
    deterministic.  window: vertices evaluated concurrently against one state (csrc/reorder.cu); 1 = the sequential
    algorithm, 0 = the library's choice."""
    e = torch.as_tensor(edge_index)
    if e.dim() != 2 or e.shape[0] != 2:
        raise RuntimeError("edge_index must have shape [2, E]")
    e = e.to(torch.int32).cpu().contiguous()
    n = int(num_nodes) if num_nodes is not None else (int(e.max()) + 1 if e.numel() else 0)
    perm = torch.empty(n, dtype=torch.int32)
    p = lambda t: ctypes.c_void_p(t.data_ptr() if t.numel() else 0)   # noqa: E731
    _lib.check(_lib.load().gnna_rabbit_reorder_host_ex(p(e[0]), p(e[1]), e.shape[1], n, p(perm), int(window)), "rabbit reorder")
    return perm


def reorder(in_edge_index):
    """Drop-in for `rabbit.reorder`: relabelled edge list, same shape/dtype/order as the input."""
    e = torch.as_tensor(in_edge_index)
    if not e.is_contiguous():
        raise RuntimeError("in_edge_index must be contiguous")        # reorder.cpp:232-233 CHECK_INPUT
    perm = permutation(e)
    out = perm[e.to(torch.int64).reshape(-1)].reshape(e.shape)
    return out.to(torch.int32)
