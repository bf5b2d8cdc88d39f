"""bench.py's contract, as far as it can be checked without a GPU: the reference arm runs on the host cores alone and
prints ONE JSON line with the keys the driver reads; the algorithmic-bytes formula is SURVEY.md 8(d)'s."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_on_cpu():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cora", "--dim", "16",
                          "--steps", "2", "--warmup", "1", "--cpu-seconds", "2"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "edge*dim/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert line["config"]["num_nodes"] == 2708 and "workload" in line["config"] and "model" not in line["config"]
    ep = line["extras"]["gcn_epoch_ms"]                     # the other half of the metric, on the same host cores
    assert ep["ms"] > 0 and ep["epochs_timed"] >= 1 and "1433-16-7" in ep["model"] and ep["final_loss"] == ep["final_loss"]
    ts = line["extras"]["torch_sparse_csr"]
    assert ts["ms"] > 0 and ts["vs_port"] > 0
    # the CPU arm must not load the product: the graph is built with torch ops, the arithmetic is the oracle's
    assert "libgnna_b200" not in out.stderr


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "cora",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_algorithmic_bytes_formula():
    sys.path.insert(0, ROOT)
    import bench
    E, N, D, P = 114615892, 232965, 64, 3696299
    # E*(D*s_x + 4) + N*(D*s_y + 8) + (2P+1)*4
    assert bench.alg_bytes(E, N, D, P) == E * (D * 4 + 4) + N * (D * 4 + 8) + (2 * P + 1) * 4
    assert bench.alg_bytes(E, N, D, P, sx=2) == E * (D * 2 + 4) + N * (D * 4 + 8) + (2 * P + 1) * 4
    assert bench.alg_bytes(E, N, D, P, gcn=True) - bench.alg_bytes(E, N, D, P) == 4 * E
    assert bench.alg_bytes(E, N, D, P, prescale=True) - bench.alg_bytes(E, N, D, P) == 2 * N * D * 4


def test_cpu_epoch_of_the_reference_arm_differentiates_the_model():
    """bench.cpu_gcn_loss_and_grads (the CPU arm's GCN epoch: oracle aggregations + torch.mm, gradients written by hand
    in the reference's operator order) against autograd on the dense closed form diag(n) A diag(n) in float64."""
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bench
    import oracle
    from gnnadvisor_osdi21_b200 import graph
    n, din, hid, cls = 300, 20, 8, 5
    rp, ci = graph.synth_graph(n, 4000, kind="rmat", seed=4)
    rpn, cin = rp.numpy(), ci.numpy()
    pp, pn = oracle.build_part(4, rpn, exact=True)
    deg = oracle.degrees(rpn)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, din, generator=g)
    y = torch.randint(0, cls, (n,), generator=g)
    w = [torch.nn.Parameter(torch.randn(din, hid, generator=g) * 0.05), torch.nn.Parameter(torch.randn(hid, cls, generator=g) * 0.05)]
    loss = bench.cpu_gcn_loss_and_grads(oracle, x, y, w, cin, pp, pn, deg, threads=2)
    A = torch.zeros(n, n, dtype=torch.float64)
    rows = np.repeat(np.arange(n), np.diff(rpn))
    A[rows, cin.astype(np.int64)] = 1.0
    dn = torch.from_numpy(deg.astype(np.float64))
    Ah = dn[:, None] * A * dn[None, :]
    w64 = [p.detach().double().requires_grad_(True) for p in w]
    out = Ah @ (torch.relu(Ah @ (x.double() @ w64[0])) @ w64[1])
    ref = torch.nn.functional.nll_loss(torch.log_softmax(out, dim=1), y)
    ref.backward()
    assert abs(loss - float(ref)) <= 1e-5 * abs(float(ref))
    for p, q in zip(w, w64):
        assert torch.allclose(p.grad.double(), q.grad, rtol=1e-3, atol=1e-4 * float(q.grad.abs().max()))
