#!/bin/bash
# Does the second pass of the streaming path hit L2?  dram bytes per kernel at several batch sizes.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
CMD="python scripts/sweep.py --batches 32,64,128,256 --iters 1 --variants stream_nochunk --out gpurun_out/sweep_l2.json"
timeout 300 $CMD > gpurun_out/l2_plain.log 2>&1 && \
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:plane_ --csv --log-file gpurun_out/l2_metrics.csv $CMD > gpurun_out/l2_ncu.log 2>&1
echo "ncu exit $?"
