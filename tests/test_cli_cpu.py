"""The command line of gnnadvisor_osdi21_b200/main.py against the reference's GNNA_main.py:15-40: the same 17 flags, types,
defaults and choices (the reference's batch drivers and log scrapers pass exactly these), plus this runtime's additions.
Parsing only -- nothing is run, no GPU."""
import os
import re

import pytest

from gnnadvisor_osdi21_b200 import main

# (flag, type, default, choices) as in GNNA_main.py:17-40
REFERENCE_FLAGS = [
    ("dataDir", str, "../osdi-ae-graphs", None), ("dataset", str, "amazon0601", None), ("dim", int, 96, None),
    ("hidden", int, 16, None), ("classes", int, 22, None), ("model", str, "gcn", ["gcn", "gin"]),
    ("num_epoches", int, 200, None), ("partSize", int, 32, None), ("dimWorker", int, 32, None),
    ("warpPerBlock", int, 4, None), ("sharedMem", int, 100, None),
    ("manual_mode", str, "True", ["True", "False"]), ("verbose_mode", str, "False", ["True", "False"]),
    ("enable_rabbit", str, "False", ["True", "False"]), ("loadFromTxt", str, "False", ["True", "False"]),
    ("single_spmm", str, "False", ["True", "False"]), ("verify_spmm", str, "False", ["True", "False"]),
]


def test_the_17_reference_flags_with_their_types_defaults_and_choices():
    actions = {a.dest: a for a in main.build_parser()._actions}
    for dest, typ, default, choices in REFERENCE_FLAGS:
        a = actions[dest]
        assert a.option_strings == ["--" + dest] and a.type is typ and a.default == default, dest
        assert (list(a.choices) if a.choices else None) == choices, dest
    args = main.build_parser().parse_args([])
    assert all(getattr(args, d) == v for d, _, v, _ in REFERENCE_FLAGS)
    # what the reference's drivers pass (0_bench_GNNA_GCN.py / s7-4_1_neighbor_partitioning.py style)
    args = main.build_parser().parse_args("--dataset citeseer --dim 3703 --hidden 16 --classes 6 --partSize 32 --model gcn "
                                          "--warpPerBlock 2 --manual_mode True --verbose_mode False --enable_rabbit True "
                                          "--loadFromTxt False --dataDir /data".split())
    assert (args.dataset, args.dim, args.enable_rabbit, args.warpPerBlock) == ("citeseer", 3703, "True", 2)
    with pytest.raises(SystemExit):
        main.build_parser().parse_args(["--manual_mode", "yes"])          # string booleans only, like the reference
    # additions of this runtime keep the reference's behaviour when absent
    assert (args.synthetic, args.fused, args.cuda_graph, args.gather_dtype, args.decider) == ("", "False", "False", "fp32", "b200")


@pytest.mark.skipif(not os.path.exists("/root/reference/GNNAdvisor/GNNA_main.py"), reason="the reference tree is only mounted in the authoring container")
def test_flag_table_is_the_reference_script_verbatim():
    text = open("/root/reference/GNNAdvisor/GNNA_main.py").read()
    found = re.findall(r"add_argument\(\s*['\"]--(\w+)['\"]\s*,\s*type=(\w+)\s*,(?:\s*choices=\[([^\]]*)\]\s*,)?\s*default=([^,]+?),(?:\s*choices=\[([^\]]*)\]\s*,)?", text)
    ref = {}
    for name, typ, ch1, default, ch2 in found:
        ch = ch1 or ch2
        ref[name] = (typ, default.strip().strip("'\""), [c.strip().strip("'\"") for c in ch.split(",")] if ch else None)
    assert len(ref) == 17
    for dest, typ, default, choices in REFERENCE_FLAGS:
        assert ref[dest] == (typ.__name__, str(default), choices), dest


REF_PY = "/root/reference/GNNAdvisor"


@pytest.mark.skipif(not os.path.isdir(REF_PY), reason="the reference tree is only mounted in the authoring container")
def test_verification_of_the_reference_run_live_agrees_with_ours(tmp_path, monkeypatch, capsys):
    """GNNA_main.py:116-125 / unitest.py:33-63: SAG on all-ones features against a CPU sparse product of the RAW edge list,
    PASSED when all but 1e-4 of the elements are EXACTLY equal.  The reference's Verification class, imported unchanged, and
    main.verify_spmm judge the same dataset object (ours) through the same extension surface (the CPU oracle here; compat/
    provides torch_sparse): both print PASSED for a correct aggregation and FAILED for a broken one."""
    import importlib
    import sys
    import numpy as np
    import torch
    import oracle
    from gnnadvisor_osdi21_b200 import graph, ops
    rng = np.random.default_rng(3)
    n, e = 300, 5000
    src, dst = rng.integers(0, n, e), rng.integers(0, n, e)
    keep = src != dst
    src, dst = np.concatenate([src[keep], dst[keep]]), np.concatenate([dst[keep], src[keep]])
    key = np.unique(src * n + dst)                           # the datasets of the reference hold every edge once
    src, dst = key // n, key % n
    npz = str(tmp_path / "g.npz")
    graph.save_npz(npz, src, dst, n)
    ds = graph.custom_dataset(npz, 16, 4, load_from_txt=False, device="cpu")
    pp, pn = (t.int() for t in ops.build_part(8, ds.row_pointers))

    class Info:
        pass
    info = Info()
    info.row_pointers, info.column_index, info.degrees = ds.row_pointers, ds.column_index, ds.degrees
    info.partPtr, info.part2Node, info.partSize, info.dimWorker, info.warpPerBlock = pp, pn, 8, 16, 4
    broken = {"on": False}

    def sag(X, rp, ci, deg, ppt, pnt, ps, dw, wpb):
        out = torch.from_numpy(oracle.SAG(X.numpy(), rp.numpy(), ci.numpy(), None, ppt.numpy(), pnt.numpy()))
        if broken["on"]:
            out[::7] += 1.0
        return out
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(ops, "SAG", sag)
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gnnadvisor_osdi21_b200", "compat")
    saved = {k: sys.modules.pop(k) for k in ("GNNAdvisor", "unitest", "torch_sparse") if k in sys.modules}
    sys.path[:0] = [compat, REF_PY]
    try:
        unitest = importlib.import_module("unitest")
        assert unitest.__file__.startswith(REF_PY)
        unitest.GNNA.SAG = sag
        for state, verdict in ((False, "PASSED"), (True, "FAILED")):
            broken["on"] = state
            capsys.readouterr()
            v = unitest.Verification(16, info.row_pointers, info.column_index, info.degrees, info.partPtr, info.part2Node, 8, 16, 4)
            v.compute()
            v.reference(ds.edge_index, ds.val, ds.num_nodes)
            v.compare()
            theirs = capsys.readouterr().out
            ok = main.verify_spmm(info, ds, 16)
            mine = capsys.readouterr().out
            assert ("# Verification " + verdict) in theirs and ("# Verification " + verdict) in mine and ok == (verdict == "PASSED")
            assert theirs.splitlines() == mine.splitlines()                  # the same three lines, in the same order
    finally:
        del sys.path[:2]
        for k in ("GNNAdvisor", "unitest", "torch_sparse"):
            sys.modules.pop(k, None)
        sys.modules.update(saved)
