"""B200-native re-implementation of the GNNAdvisor (OSDI'21) aggregation hot path.

Public surface (the reference's extension API, GNNAdvisor/GNNConv/GNNAdvisor.cpp:253-263):
    SAG, forward, backward, forward_gin, backward_gin, build_part
plus the host-side mirror of the reference's Python layer (layers.py = gnn_conv.py,
param.py = param.py) and graph helpers.  The compute lives in libgnna_b200.so (csrc/, C ABI in
include/gnna_b200.h); importing this package never falls back to a CPU/torch implementation.
"""
from .ops import (SAG, forward, backward, forward_gin, backward_gin, build_part,  # noqa: F401
                  build_part_exact, aggregate_bf16, degrees_from_row_ptr, launch_info)

__version__ = "0.1.0"
