"""ONE launch of every kernel of libgnna_b200.so on a look-alike graph -- the process `ncu --set full` attaches to
(tools/ncu_set.sh).  python tools/ncu_set.py [reddit|ogbn-products] [D]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from gnnadvisor_osdi21_b200 import _lib, graph, ops  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "reddit"
D = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda:0")
lib = _lib.load()
gr = graph.lookalike(wl, device=dev)
rp, ci = gr["row_ptr"], gr["col_idx"]
N = gr["num_nodes"]
torch.cuda.synchronize()
print("graph ready", wl, N, ci.numel(), flush=True)
pp, pn = ops.build_part(32, rp)                                   # part_tile_sums / scan / part_offsets / part_expand
deg = ops.degrees_from_row_ptr(rp)                                # degrees_kernel
X = torch.randn(N, D, device=dev)
W = torch.eye(D, device=dev)
a = (rp, ci, deg, pp, pn, 32, 32, 4)
ops.forward(X, W, *a)                                             # SGEMM + repack_rows (pre-scale) + aggregate_kernel fp32
ops.forward_gin(X, W, rp, ci, 0.5, pp, pn, 32, 32, 4)             # aggregate_kernel fp32 (GIN flags)
if wl == "reddit":                                                # the layer-1 products: gemm_tf32x3_kernel NN (X*W) and TN (X^T*G)
    X602 = torch.randn(N, 602, device=dev)
    W602 = torch.randn(602, D, device=dev) / 8
    ops.forward(X602, W602, *a)
    ops.backward(X, X602, W602, *a, need_d_input=False)
    del X602
Xb = ops.scale_rows_bf16(X, deg)                                  # scale_rows_bf16_kernel
ops.aggregate_bf16(3, Xb, rp, ci, deg, 1.0, pp, pn, 32, 32, 4)    # aggregate_runs (bf16) by the library's rule
prev = _lib.set_runs(0)
ops.aggregate_bf16(3, Xb, rp, ci, deg, 1.0, pp, pn, 32, 32, 4)    # aggregate_kernel bf16
_lib.set_runs(4)
ops.SAG(X, *a)                                                    # aggregate_runs fp32
_lib.set_runs(prev)
_lib.set_staged(True)
ops.SAG(X, *a)                                                    # aggregate_staged (TMA ring)
_lib.set_staged(False)
ops.aggregate_gemm_fused(2, Xb, W, rp, ci, None, 0.5, pp, pn, 32, 32, 4, want_agg=True)   # fused tcgen05 tile, bf16 rows
X41 = torch.randn(N, 41, device=dev)
ops.SAG(X41, *a)                                                  # repack (pad) + aggregate + unpack_rows
# launch-bound graph: the single-launch kernel
c = graph.lookalike("cora", device=dev)
cpp, cpn = ops.build_part(32, c["row_ptr"])
cdeg = ops.degrees_from_row_ptr(c["row_ptr"])
Xc = torch.randn(c["num_nodes"], 16, device=dev)
ops.forward(Xc, torch.eye(16, device=dev), c["row_ptr"], c["col_idx"], cdeg, cpp, cpn, 32, 16, 8)   # aggregate_small_kernel
# the L2 probe (bench.py's roof)
buf = torch.zeros(30 << 18, device=dev)
sink = torch.zeros(1, dtype=torch.int32, device=dev)
per = ctypes.c_int64(0)
for mode in (0, 0, 1):
    _lib.check(lib.gnna_probe_l2_read(ctypes.c_void_p(buf.data_ptr()), 30 << 20, 8, mode, 1536, 128, ctypes.c_void_p(sink.data_ptr()),
                                      ctypes.byref(per), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "probe")
torch.cuda.synchronize()
print("done", wl, D)
