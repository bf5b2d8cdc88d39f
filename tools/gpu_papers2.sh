#!/bin/bash
# config #5 at reduced scale on 2 GPUs, row-weight study: gpurun --gpus 2 -- bash tools/gpu_papers2.sh
mkdir -p gpurun_out
for w in 0 8; do
GNNA_ROW_WEIGHT=$w timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$w bench.py --gpus 2 --workload ogbn-papers100M --scale 0.05 --dim 128 --steps 20 --warmup 3 > gpurun_out/config5_papers_x0.05_2gpu_w$w.json 2> gpurun_out/config5_papers_2gpu_w$w.err; echo "w=$w exit=$?"; python -c "
import json; c=json.load(open('gpurun_out/config5_papers_x0.05_2gpu_w$w.json')); print(c['ms_per_step'], c['extras']['ms_kernel_only'], c['extras']['ms_exchange_only'], [(s['edges'], s['rows'], s['halo_rows']) for s in c['extras']['shards']])"
done
