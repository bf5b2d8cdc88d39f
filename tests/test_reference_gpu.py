"""The reference ITSELF on the GPU box, beside this runtime:

  * its own CUDA kernels (oracle/_ref/GNNAdvisor_ref.so, compiled from /root/reference by oracle/build_ref.py) called live
    on a 0.25x Reddit look-alike -- forward / backward / forward_gin / backward_gin / SAG of the product must agree with
    them within BASELINE.json's 1e-4 (what bench.py's ref_gpu leg only printed in round 1);
  * its own UNCHANGED scripts (GNNA_main.py, gnn_conv.py, param.py, dataset.py, unitest.py -- staged byte for byte in
    oracle/_ref/ref_py.zip) run as a subprocess with gnnadvisor_osdi21_b200/compat on PYTHONPATH: training (GCN, GIN),
    --verify_spmm and --single_spmm, i.e. the drop-in promise of SURVEY.md 8(b) executed on a GPU;
  * the same unchanged script on the reference's own kernels, for the epoch time the reference gets on this box.

Skipped when the artefacts were not built (they are built by __graft_entry__.build() in the authoring container and
travel to the GPU box with the snapshot; /root/reference is never read here)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import build_ref
import oracle
from helpers import assert_close
from gnnadvisor_osdi21_b200 import graph, ops

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "gnnadvisor_osdi21_b200", "compat")


@pytest.fixture(scope="module")
def ref():
    m = build_ref.load_ref()
    if m is None:
        pytest.skip("oracle/_ref/GNNAdvisor_ref.so not built")
    return m


@pytest.fixture(scope="module")
def quarter_reddit():
    dev = torch.device("cuda:0")
    gr = graph.lookalike("reddit", device=dev, scale=0.25)
    rp, ci = gr["row_ptr"], gr["col_idx"]
    pp, pn = ops.build_part(32, rp)
    deg = ops.degrees_from_row_ptr(rp)
    return gr["num_nodes"], rp, ci, pp, pn, deg


def _rel(got, ref_t, terms=None):
    """max |got - ref| / max(|ref|, 1e-3 ||ref||_inf [, 0.1 * terms]) on the device (the matrices are 58 K x 64 ... x 602)."""
    scale = torch.maximum(ref_t.abs(), 1e-3 * ref_t.abs().max())
    if terms is not None:
        scale = torch.maximum(scale, 0.1 * terms)
    return float(((got - ref_t).abs() / scale.clamp_min(1e-30)).max())


@pytest.mark.parametrize("din,dout,ps,dw,wpb", [(64, 64, 32, 32, 4), (96, 41, 32, 32, 8), (100, 64, 16, 16, 2)])
def test_product_agrees_with_the_reference_kernels_live(ref, quarter_reddit, din, dout, ps, dw, wpb):
    n, rp, ci, pp32, pn32, deg = quarter_reddit
    dev = rp.device
    pp, pn = (pp32, pn32) if ps == 32 else ops.build_part(ps, rp)
    gen = torch.Generator(device=dev).manual_seed(1)
    X = torch.randn(n, din, device=dev, generator=gen)
    W = (torch.rand(din, dout, device=dev, generator=gen) * 2 - 1) / dout ** 0.5
    dO = torch.randn(n, dout, device=dev, generator=gen)
    a = (rp, ci, deg, pp, pn, ps, dw, wpb)
    g = (rp, ci, 0.5, pp, pn, ps, dw, wpb)
    # sums of |terms| bound what any summation order may do to cancelling elements (helpers.assert_close)
    sag_abs = ops.SAG(X.abs(), *a)
    assert _rel(ops.SAG(X, *a), ref.SAG(X, *a), sag_abs) <= 1e-4
    T_abs = ops.forward(X.abs(), W.abs(), *a)[0]
    assert _rel(ops.forward(X, W, *a)[0], ref.forward(X, W, *a)[0], T_abs) <= 1e-4
    dX, dW = ops.backward(dO, X, W, *a)
    rdX, rdW = ref.backward(dO, X, W, *a)
    adX, adW = ops.backward(dO.abs(), X.abs(), W.abs(), *a)
    assert _rel(dX, rdX, adX) <= 1e-4 and _rel(dW, rdW, adW) <= 1e-4
    o, S = ops.forward_gin(X, W, *g)
    ro, rS = ref.forward_gin(X, W, *g)
    ao, aS = ops.forward_gin(X.abs(), W.abs(), *g)
    assert _rel(S, rS, aS) <= 1e-4 and _rel(o, ro, ao) <= 1e-4
    dXg, dWg = ops.backward_gin(dO, rS, W, *g)
    rdXg, rdWg = ref.backward_gin(dO, rS, W, *g)
    adXg, adWg = ops.backward_gin(dO.abs(), rS.abs(), W.abs(), *g)
    assert _rel(dXg, rdXg, adXg) <= 1e-4 and _rel(dWg, rdWg, adWg) <= 1e-4


# ------------------------------------------------------------------------------------------ the unchanged scripts
@pytest.fixture(scope="module")
def ref_scripts(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("refpy"))
    got = build_ref.unpack_py(d)
    if got is None:
        pytest.skip("oracle/_ref/ref_py.zip not staged")
    return got


@pytest.fixture(scope="module")
def small_dataset(tmp_path_factory):
    """An amazon0505-shaped file in the reference's .npz format (dataset.py:87-91): 41 K nodes, ~0.49 M edges."""
    d = str(tmp_path_factory.mktemp("graphs"))
    rp, ci = graph.synth_graph(41023, 487886, kind="rmat", seed=3)
    rows = torch.repeat_interleave(torch.arange(rp.numel() - 1), (rp[1:] - rp[:-1]).long())
    graph.save_npz(os.path.join(d, "small.npz"), rows.numpy(), ci.numpy(), rp.numel() - 1)
    return d


def _run_main(py_dir, module_dir, data_dir, extra, timeout=600):
    env = dict(os.environ)
    # `GNNAdvisor` resolves to module_dir (this runtime's compat/ or the compiled reference); dgl / rabbit / torch_sparse
    # always to the stand-ins in compat/ (none of them is installable offline)
    env["PYTHONPATH"] = os.pathsep.join([module_dir, COMPAT, py_dir, os.path.join(ROOT, "oracle"), env.get("PYTHONPATH", "")])
    cmd = [sys.executable, os.path.join(py_dir, "GNNA_main.py"), "--dataDir", data_dir, "--dataset", "small", "--dim", "96",
           "--hidden", "16", "--classes", "22"] + extra
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=timeout, cwd=py_dir)
    return out


def _time_ms(stdout):
    m = re.search(r"Time \(ms\): (\d+\.\d+)", stdout)          # GNNA_main.py:202
    return float(m.group(1)) if m else None


@pytest.mark.parametrize("model,wpb", [("gcn", "8"), ("gin", "2")])
def test_unchanged_GNNA_main_trains_on_this_runtime(ref_scripts, small_dataset, model, wpb):
    py_dir, refmod = ref_scripts
    ours = _run_main(py_dir, COMPAT, small_dataset, ["--model", model, "--num_epoches", "20", "--warpPerBlock", wpb])
    assert ours.returncode == 0, ours.stderr[-3000:]
    t_ours = _time_ms(ours.stdout)
    assert t_ours is not None and t_ours > 0, ours.stdout[-2000:]
    line = "[unchanged GNNA_main.py, %s, amazon0505/10 look-alike] this runtime: %.3f ms/epoch" % (model, t_ours)
    if os.path.exists(build_ref.SO):
        theirs = _run_main(py_dir, refmod, small_dataset, ["--model", model, "--num_epoches", "20", "--warpPerBlock", wpb])
        assert theirs.returncode == 0, theirs.stderr[-3000:]
        line += "; the reference's own kernels: %.3f ms/epoch" % _time_ms(theirs.stdout)
    print(line)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "unchanged_GNNA_main.txt"), "a") as f:
        f.write(line + "\n")


def test_unchanged_GNNA_main_verify_and_single_spmm(ref_scripts, small_dataset):
    py_dir, _ = ref_scripts
    v = _run_main(py_dir, COMPAT, small_dataset, ["--verify_spmm", "True"])
    assert v.returncode == 0 and "# Verification PASSED" in v.stdout, (v.stdout[-2000:], v.stderr[-2000:])   # unitest.py:58-63
    s = _run_main(py_dir, COMPAT, small_dataset, ["--single_spmm", "True", "--num_epoches", "20"])
    assert s.returncode == 0 and re.search(r"SpMM profiling avg \(ms\): \d+\.\d+", s.stdout), (s.stdout[-2000:], s.stderr[-2000:])


def test_unchanged_GNNA_main_auto_mode_with_rabbit(ref_scripts, small_dataset):
    py_dir, _ = ref_scripts
    r = _run_main(py_dir, COMPAT, small_dataset, ["--manual_mode", "False", "--enable_rabbit", "True", "--verbose_mode", "True",
                                                  "--num_epoches", "5"])
    assert r.returncode == 0 and _time_ms(r.stdout) is not None, (r.stdout[-2000:], r.stderr[-3000:])
