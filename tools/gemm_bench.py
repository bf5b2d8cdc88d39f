"""The two tall-skinny products of the Reddit GCN layer 1 (X*W: 232 965 x 602 x 64, X^T*G: 602 x 232 965 x 64) through
gnna_sgemm_f32: cuBLAS SGEMM (mode 0) against the tcgen05 3xTF32 kernels (1 lockstep, 2 warp-specialised).
CUDA events, 20 launches after 3 warm-ups; error against a float64 product.   python tools/gemm_bench.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from gnnadvisor_osdi21_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
p = lambda t: ctypes.c_void_p(t.data_ptr())   # noqa: E731


def run(ta, A, B, m, n, k):
    C = torch.empty(m, n, device=dev)
    f = lambda: _lib.check(lib.gnna_sgemm_f32(int(ta), 0, m, n, k, p(A), p(B), p(C), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "sgemm")   # noqa: E731
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(20):
        f()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / 20, C


for N, K, D in ((232965, 602, 64), (232965, 602, 41), (2449029, 100, 64)):
    X = torch.randn(N, K, device=dev)
    W = torch.randn(K, D, device=dev) / K ** 0.5
    G = torch.randn(N, D, device=dev)
    ref_nn = (X[:4096].double() @ W.double())
    terms_nn = X[:4096].abs().double() @ W.abs().double()
    for mode in (0, 1, 2):
        _lib.set_tc_gemm(mode)
        ms, C = run(False, X, W, N, D, K)
        err = float(((C[:4096].double() - ref_nn).abs() / terms_nn).max())
        ms2, C2 = run(True, X, G, K, D, N)
        print("N=%d K=%d D=%d mode %d:  X*W %.3f ms (%.0f GB/s of X, err/terms %.1e)   X^T*G %.3f ms (%.0f GB/s)"
              % (N, K, D, mode, ms, N * K * 4 / ms / 1e6, err, ms2, N * K * 4 / ms2 / 1e6), flush=True)
    del X, W, G
