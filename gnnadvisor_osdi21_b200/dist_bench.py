"""N > 1 arm of bench.py: the same aggregation step on a graph sharded over N GPUs of one node.

One rank per GPU (torchrun), NCCL.  The graph is FIXED (the Reddit look-alike of the N=1 run), so this
is strong scaling: rank r owns a contiguous vertex range holding ~E/N edges (dist.ShardedGraph), one
step = halo exchange of the remote neighbour rows (gather + ONE all_to_all_single over NVLink) + the
local aggregation kernel.  value = E_global * D / max-over-ranks device time.
"""
import json
import os

import torch
import torch.distributed as dist


def run(args, bench):
    from . import _lib, graph, dist as gdist
    # stdout carries exactly ONE JSON line (rank 0): anything libraries print there (NCCL's version
    # banner does) is sent to stderr instead
    import sys
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)

    # every rank builds the same seeded graph on its own GPU, keeps its shard, drops the rest
    gr = graph.lookalike(args.workload, device=device, scale=args.scale)
    rp, ci = gr["row_ptr"], gr["col_idx"]
    N, E, D = gr["num_nodes"], ci.numel(), args.dim
    row_weight = int(os.environ.get("GNNA_ROW_WEIGHT", gdist.default_row_weight(world)))
    sg = gdist.ShardedGraph(rp, ci, args.part_size, device=device, row_weight=row_weight).build_tables()
    gen = torch.Generator(device=device).manual_seed(20212)
    X = torch.randn(N, D, device=device, generator=gen)
    x_ext = sg.new_features(D)
    sg.local(x_ext).copy_(X[sg.v0:sg.v0 + sg.n_local])
    del X, rp, ci, gr["row_ptr"], gr["col_idx"]
    torch.cuda.empty_cache()
    out = torch.empty(sg.n_local, D, device=device)
    P_local = sg.part2node.numel()

    # halo exchange: NVLink push kernel over CUDA-IPC mapped peer buffers (csrc/halo.cu); NCCL
    # all_to_all_single if the box does not allow IPC mappings
    peer, halo_mode = None, "nccl all_to_all_single"
    if os.environ.get("GNNA_HALO", "peer") == "peer":
        try:
            peer = gdist.PeerHalo(sg, D)
            for b in (1, 2):
                peer.features(b).copy_(x_ext)
            halo_mode = "NVLink push kernel over CUDA IPC (gnna_halo_push_f32)"
        except Exception as e:   # noqa: BLE001
            peer, halo_mode = None, "nccl all_to_all_single (peer mapping failed: %s)" % str(e)[:80]
    ok = torch.tensor([1 if peer is not None else 0], device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 0 and peer is not None:
        peer = None
        halo_mode = "nccl all_to_all_single (peer mapping failed on another rank)"

    def step_serial():
        sg.aggregate(1, x_ext, out=out, dim_worker=args.dim_worker, warp_per_block=args.warp_per_block, peer=peer)

    # overlapped step: the rank's rows are written pre-scaled into the step buffer (what the X*W epilogue of a
    # layer would do), the push runs on a second stream, per-owner sub-shards are aggregated as their rows land
    overlap = peer is not None and D % 4 == 0 and os.environ.get("GNNA_OVERLAP", "1") == "1"
    x_src = sg.local(x_ext).clone() if overlap else None
    if overlap:
        sg.build_owner_shards()

    def step_overlap():
        sg.write_local(peer, x_src, prescale=True)
        sg.aggregate_overlapped(1, peer, out, dim_worker=args.dim_worker, warp_per_block=args.warp_per_block)

    step = step_overlap if overlap else step_serial

    # the overlapped step is ~12 small launches on two streams: replay it from CUDA graphs (one per buffer
    # parity; the step counter lives in device memory) so the host is not the bottleneck at 8 GPUs
    graphs, launches_per_step = None, None
    if overlap and os.environ.get("GNNA_GRAPH", "1") == "1":
        try:
            for _ in range(3):
                step_overlap()
            torch.cuda.synchronize()
            dist.barrier(device_ids=[local])
            base_step = peer.step
            graphs = []
            for i in range(2):
                g = torch.cuda.CUDAGraph()
                c0 = _lib.launch_count()
                with torch.cuda.graph(g):
                    step_overlap()
                launches_per_step = _lib.launch_count() - c0
                graphs.append(g)
            peer.step = base_step                       # nothing ran during capture
            state = {"i": 0}

            def step_graph():
                graphs[state["i"] & 1].replay()
                state["i"] += 1
                peer.step += 1
            step = step_graph
        except Exception as e:   # noqa: BLE001
            graphs = None
            halo_mode += " (graph capture failed: %s)" % str(e)[:60]
            step = step_overlap

    halo_check, halo_diff_elems = None, None
    if peer is not None:   # the two exchange implementations must give the same aggregation
        ref_out = sg.aggregate(1, x_ext, dim_worker=args.dim_worker, warp_per_block=args.warp_per_block).clone()
        step_serial()
        step()
        step()
        step()
        halo_check = ((out - ref_out).abs().max() / ref_out.abs().max().clamp_min(1e-30)).item()
        halo_diff_elems = int((out != ref_out).sum().item())
        out.zero_()                                    # and once more from a zeroed output: the step must rewrite it
        step()
        torch.cuda.synchronize()
        halo_check = max(halo_check, ((out - ref_out).abs().max() / ref_out.abs().max().clamp_min(1e-30)).item())
        assert halo_check < 1e-4, "overlapped / peer exchange disagrees with the NCCL path: %g" % halo_check

    def kernel_only():
        sg.aggregate(1, x_ext, out=out, dim_worker=args.dim_worker, warp_per_block=args.warp_per_block, do_exchange=False)

    def exchange_only():
        if peer is not None:
            peer.exchange()
            peer.ack()
        else:
            sg.exchange(x_ext)

    def reduce_max(ms):
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    barrier = lambda: dist.barrier(device_ids=[local])   # noqa: E731
    sampler = bench.ClockSampler(local) if rank == 0 else None
    step(); torch.cuda.synchronize()
    _lib.launch_count(reset=True)
    if sampler:
        sampler.start()
    ms = reduce_max(bench.timed(step, args.steps, args.warmup, barrier)) / args.steps
    clocks = sampler.stop() if sampler else None
    launches = _lib.launch_count() * args.steps // (args.steps + args.warmup)
    if graphs is not None:
        launches = launches_per_step * args.steps
    k = max(3, args.steps // 4)
    ms_kernel = reduce_max(bench.timed(kernel_only, k, 3, barrier)) / k
    ms_exch = reduce_max(bench.timed(exchange_only, k, 3, barrier)) / k
    ms_serial = reduce_max(bench.timed(step_serial, k, 3, barrier)) / k if overlap else ms

    # end to end: this rank's features start and end in pinned host memory; H2D of step i+1 and D2H of
    # step i-1 overlap the exchange + aggregation of step i (three streams, double-buffered staging)
    x_host = sg.local(x_ext).cpu().pin_memory()
    out_host = [torch.empty(sg.n_local, D).pin_memory() for _ in range(2)]
    x_stage = [torch.empty(sg.n_local, D, device=device) for _ in range(2)]
    o_stage = [torch.empty(sg.n_local, D, device=device) for _ in range(2)]
    s_in, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_used = [torch.cuda.Event() for _ in range(2)]
    ev_o = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    cnt = {"i": 0}

    def e2e_step():
        i = cnt["i"] & 1
        first = cnt["i"] < 2
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(s_in):
            if not first:
                s_in.wait_event(ev_used[i])
            x_stage[i].copy_(x_host, non_blocking=True)
            ev_in[i].record(s_in)
        cur.wait_event(ev_in[i])
        if overlap:
            x_src.copy_(x_stage[i])
        else:
            sg.local(peer.features() if peer is not None else x_ext).copy_(x_stage[i])
        ev_used[i].record(cur)
        step()
        if not first:
            cur.wait_event(ev_done[i])
        o_stage[i].copy_(out)
        ev_o[i].record(cur)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_o[i])
            out_host[i].copy_(o_stage[i], non_blocking=True)
            ev_done[i].record(s_out)
        cnt["i"] += 1

    ke = max(4, min(args.steps, 30))
    for _ in range(4):
        e2e_step()
    torch.cuda.synchronize()
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    s_in.wait_event(t0)
    for _ in range(ke):
        e2e_step()
    torch.cuda.current_stream().wait_event(ev_done[0])
    torch.cuda.current_stream().wait_event(ev_done[1])
    t1.record()
    torch.cuda.synchronize()
    barrier()
    ms_e2e = reduce_max(t0.elapsed_time(t1)) / ke

    halo = sg.halo_bytes(D)
    stats = torch.tensor([sg.num_edges_local, sg.n_local, sg.n_halo, P_local, halo["recv"], halo["send"]],
                         device=device, dtype=torch.float64)
    allstats = [torch.zeros_like(stats) for _ in range(world)]
    dist.all_gather(allstats, stats)
    if rank == 0:
        peak, peak_src = bench.peak_gbs()
        # roofline of the local kernel on the most loaded rank
        per_rank = [[float(v) for v in s.cpu()] for s in allstats]
        worst = max(per_rank, key=lambda s: s[0])
        B = bench.alg_bytes(int(worst[0]), int(worst[1]), D, int(worst[3]))
        achieved = B / (ms_kernel * 1e-3) / 1e9
        line = {"metric": "aggregation throughput (GCN SpMM), edges*dim/s", "value": E * D / (ms * 1e-3), "unit": "edge*dim/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": bench.config_of(args, N, E, int(sum(s[3] for s in per_rank)),
                                          {"parallelism": "1-D vertex-range shards x%d (cost-balanced: edges + %d per row), one halo exchange per step" % (world, row_weight),
                                           "halo_exchange": halo_mode + ("; per-owner sub-shards aggregated while the rest is in flight" if overlap else "")}),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": None, "kernel": "gnna::aggregate_kernel<float,4,16,1,false> on the most loaded shard",
                             "alg_bytes_per_launch": B, "peak_source": peak_src,
                             "note": "per-GPU kernel, exchange excluded; max over ranks of the kernel-only time"},
                "e2e": {"value": E * D / (ms_e2e * 1e-3), "unit": "edge*dim/s", "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": int(sum(s[1] for s in per_rank) * D * 4),
                        "d2h_bytes_per_step": int(sum(s[1] for s in per_rank) * D * 4)},
                "gpu_launches": int(launches), "clocks": clocks, "impl": "ours",
                "extras": {"ms_kernel_only": ms_kernel, "ms_exchange_only": ms_exch, "ms_step_without_overlap": ms_serial,
                           "overlap": bool(overlap), "cuda_graph": graphs is not None,
                           "peer_vs_nccl_max_rel_diff": halo_check, "peer_vs_nccl_differing_elements": halo_diff_elems,
                           "shards": [{"edges": int(s[0]), "rows": int(s[1]), "halo_rows": int(s[2]),
                                       "halo_recv_bytes": int(s[4]), "halo_send_bytes": int(s[5])} for s in per_rank]}}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if peer is not None:
        err = peer.error()
        peer.close()
        assert err == 0, "halo wait timed out (%d)" % err
    dist.barrier(device_ids=[local])
    dist.destroy_process_group()
