"""A/B of two builds of libgnna_b200.so on the same device tensors, in ONE process (run on the GPU box).

    python tools/ab_chain.py [--out gpurun_out/ab_chain.json] [--workloads reddit,ogbn-products] [--scale 1.0]

A = the in-tree default library; B, C, D = variant builds next to it (tools/build_variants.sh):
    B libgnna_b200_nochain.so  -DGNNA_CHAIN=0        the kernels before the dependent-load chain was shortened
    C libgnna_b200_hoist.so    -DGNNA_CHAIN_IDS=0    flush loads hoisted only
    D libgnna_b200_ids.so      -DGNNA_CHAIN_HOIST=0  32 ids per round trip only
    E libgnna_b200_bf16n.so    -DGNNA_BF16_NARROW=1  128-bit bf16 rows under the 3-CTA / 40-register budget
    A_runsR                    the default build after gnna_set_runs(R): run-based pipelined kernel (aggregate_runs.cu)
All are called through the C ABI (include/gnna_b200.h) with the same pointers; results are compared element-wise
with B, then each call is timed with CUDA events (3 warm-ups + 20 calls, variants interleaved twice, best kept)."""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from gnnadvisor_osdi21_b200 import _lib, graph, ops  # noqa: E402


def bind(path):
    lib = ctypes.CDLL(path)
    for name in ("gnna_sag_f32", "gnna_gcn_aggregate_f32", "gnna_gin_aggregate_f32", "gnna_aggregate_bf16",
                 "gnna_aggregate_gemm_fused_bf16", "gnna_set_runs"):
        res, args = _lib.SIGNATURES[name]
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    lib.gnna_last_error.restype = ctypes.c_char_p
    return lib


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ab_chain.json"))
    ap.add_argument("--workloads", default="reddit,ogbn-products")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--runs", default="-1,4,8", help="run lengths of the run-based kernel to time (gnna_set_runs)")
    ap.add_argument("--extra-libs", default="", help="name=path,... more variant builds")
    ap.add_argument("--dims", default="16,32,48,64,128")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    pkg = os.path.join(ROOT, "gnnadvisor_osdi21_b200")
    paths = {"A_default": _lib.LIB_PATH, "B_nochain": os.path.join(pkg, "libgnna_b200_nochain.so"),
             "C_hoist": os.path.join(pkg, "libgnna_b200_hoist.so"), "D_ids": os.path.join(pkg, "libgnna_b200_ids.so"),
             "E_bf16n": os.path.join(pkg, "libgnna_b200_bf16n.so")}
    paths.update({k: v for k, v in (kv.split("=", 1) for kv in args.extra_libs.split(",") if kv)})
    libs = {k: bind(v) for k, v in paths.items() if os.path.exists(v)}
    assert "A_default" in libs and "B_nochain" in libs, "build the variants first (tools/build_variants.sh)"
    # the run-based kernel (csrc/aggregate_runs.cu) is a run-time switch of the default build: pseudo-variants of A
    runs_of = {k: 0 for k in libs}
    for r in (int(v) for v in args.runs.split(",") if v):
        name = "A_runs%d" % r if r > 0 else "A_auto"      # -1: the library's own rule
        libs[name] = libs["A_default"]
        runs_of[name] = r
    p = lambda t: ctypes.c_void_p(t.data_ptr())   # noqa: E731
    rows = []
    for wl in args.workloads.split(","):
        gr = graph.lookalike(wl, device=dev, scale=args.scale)
        rp, ci = gr["row_ptr"], gr["col_idx"]
        pp, pn = ops.build_part(32, rp)
        deg = ops.degrees_from_row_ptr(rp)
        N, E, P = gr["num_nodes"], ci.numel(), pn.numel()
        for D in (int(v) for v in args.dims.split(",")):
            X = torch.randn(N, D, device=dev)
            Xb = X.to(torch.bfloat16)
            outs = {k: torch.empty(N, D, device=dev) for k in libs}
            st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            cases = {
                "sag_f32": lambda lib, o: lib.gnna_sag_f32(p(X), p(o), p(rp), p(ci), p(pp), p(pn), N, D, P, 32, 32, 4, st),
                "gcn_f32": lambda lib, o: lib.gnna_gcn_aggregate_f32(p(X), p(o), p(rp), p(ci), p(deg), p(pp), p(pn), N, D, P, 32, 32, 4, st),
                "gin_f32_dw_auto": lambda lib, o: lib.gnna_gin_aggregate_f32(p(X), p(o), p(rp), p(ci), 0.5, p(pp), p(pn), N, D, P, 32, 0, 4, st),
                "sag_bf16": lambda lib, o: lib.gnna_aggregate_bf16(0, p(Xb), p(o), p(rp), p(ci), p(deg), 1.0, p(pp), p(pn), N, D, P, 32, 32, 4, st),
                "gcn3_bf16": lambda lib, o: lib.gnna_aggregate_bf16(3, p(Xb), p(o), p(rp), p(ci), p(deg), 1.0, p(pp), p(pn), N, D, P, 32, 32, 4, st),
                "gcn3_bf16_dw4": lambda lib, o: lib.gnna_aggregate_bf16(3, p(Xb), p(o), p(rp), p(ci), p(deg), 1.0, p(pp), p(pn), N, D, P, 32, 4, 4, st),
            }
            if D in (64, 128):   # fused aggregate -> X*W tile (tcgen05), GIN form, dout = D; x_agg written
                Wf = (torch.rand(D, D, device=dev) * 2 - 1) / D ** 0.5
                xaggs = {k: torch.empty(N, D, device=dev) for k in libs}
                nul = ctypes.c_void_p(0)
                cases["fused_gin_f32x"] = lambda lib, o: lib.gnna_aggregate_gemm_fused_bf16(
                    2, p(X), 0, p(Wf), 0.5, p(o), nul, p(rp), p(ci), nul, p(pp), p(pn), N, D, D, P, 32, 32, 4, st)
                cases["fused_gin_bf16x"] = lambda lib, o: lib.gnna_aggregate_gemm_fused_bf16(
                    2, p(Xb), 1, p(Wf), 0.5, p(o), nul, p(rp), p(ci), nul, p(pp), p(pn), N, D, D, P, 32, 32, 4, st)
                cases["fused_gin_bf16x_xagg"] = lambda lib, o: lib.gnna_aggregate_gemm_fused_bf16(
                    2, p(Xb), 1, p(Wf), 0.5, p(o), p(xaggs["A_default"]), p(rp), p(ci), nul, p(pp), p(pn), N, D, D, P, 32, 32, 4, st)
            for cname, call in cases.items():
                row = {"workload": wl, "N": N, "E": E, "D": D, "case": cname}
                # the fused tile has two versions only (GNNA_CHAIN): C and D contain A's
                use = {k: v for k, v in libs.items() if not cname.startswith("fused") or k in ("A_default", "B_nochain")}
                for k in use:
                    if k not in outs:
                        outs[k] = torch.empty(N, D, device=dev)
                for k, lib in use.items():
                    lib.gnna_set_runs(runs_of[k])
                    rc = call(lib, outs[k])
                    if rc != 0:
                        row[k + "_error"] = (lib.gnna_last_error() or b"?").decode()
                torch.cuda.synchronize()
                b = outs["B_nochain"]
                row["max_rel_diff_vs_B"] = max(((outs[k] - b).abs().max() / b.abs().max().clamp_min(1e-30)).item() for k in use)
                best = {k: 1e30 for k in use}
                for _ in range(2):
                    for k, lib in use.items():
                        lib.gnna_set_runs(runs_of[k])
                        best[k] = min(best[k], timed(lambda: call(lib, outs[k])))
                        lib.gnna_set_runs(0)
                for k in use:
                    row[k + "_ms"] = round(best[k], 4)
                row["speedup_A_over_B"] = round(best["B_nochain"] / best["A_default"], 3)
                row["best"] = min(best, key=best.get)
                rows.append(row)
                print(row, flush=True)
            del X, Xb, outs
        del gr, rp, ci, pp, pn, deg
        torch.cuda.empty_cache()
    libs["A_default"].gnna_set_runs(-1)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"libs": {k: paths.get(k, paths["A_default"] + " + gnna_set_runs(%d)" % runs_of[k]) for k in libs}, "rows": rows},
              open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
