#!/bin/bash
# `ncu --set full` of ONE launch of every kernel of the library on a look-alike graph (one GPU); the .ncu-rep stays on the
# box, the raw-page CSV comes back in gpurun_out/ and is summarised by tools/ncu_summary.py into profiles/.
#   gpurun --timeout 1500 -- bash tools/ncu_set.sh reddit      (or ogbn-products)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WL=${1:-reddit}
REP=/tmp/ncu_set_$WL
timeout 1400 ncu --set full --clock-control none --import-source on -f -o "$REP" \
    -k regex:'aggregate|part_|scan_|repack|scale_rows|unpack_rows|degrees_kernel|l2_read|gemm_tf32|stream_pairs' -c 60 python tools/ncu_set.py "$WL" 64 > "gpurun_out/ncu_set_$WL.log" 2>&1
echo "ncu rc=$?"; tail -3 "gpurun_out/ncu_set_$WL.log"
ncu -i "$REP.ncu-rep" --page raw --csv > "gpurun_out/ncu_set_$WL.raw.csv" 2>/dev/null
python tools/ncu_summary.py "gpurun_out/ncu_set_$WL.raw.csv" "ncu --set full --clock-control none, one launch of every kernel, $WL look-alike D=64 (tools/ncu_set.sh)" > "gpurun_out/ncu_set_$WL.txt"
wc -l "gpurun_out/ncu_set_$WL.raw.csv" "gpurun_out/ncu_set_$WL.txt"
