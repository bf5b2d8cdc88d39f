#!/bin/bash
# Variant builds of libgnna_b200.so for tools/ab_chain.py (git-ignored *.so; they travel to the GPU box with the snapshot)
set -e
cd "$(dirname "$0")/.."
P=gnnadvisor_osdi21_b200
python -m $P.build -DGNNA_CHAIN=0 --out=$PWD/$P/libgnna_b200_nochain.so &
python -m $P.build -DGNNA_CHAIN_IDS=0 --out=$PWD/$P/libgnna_b200_hoist.so &
python -m $P.build -DGNNA_CHAIN_HOIST=0 --out=$PWD/$P/libgnna_b200_ids.so &
python -m $P.build -DGNNA_BF16_NARROW=1 --out=$PWD/$P/libgnna_b200_bf16n.so &
wait
ls -la $P/*.so
