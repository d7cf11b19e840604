#!/bin/bash
# Round-2 ncu evidence, run on the GPU box from the repo root:  bash scripts/ncu_r2.sh
# (each target is first run plain; ncu only profiles a command that exited 0 without it)
set -u
OUT=gpurun_out
mkdir -p $OUT
for shape in "128 28" "256 14" "512 7"; do
  set -- $shape
  python scripts/tile_one.py --path auto --c $1 --h $2 --n 1024 --bwd --reps 1 > $OUT/plain_$1.log 2>&1 || { echo "plain run failed for $1"; continue; }
  ncu --set full --clock-control none --import-source on -k regex:'tile_pipeline|l2_|resident|fused' \
      -o $OUT/r2_block_$1 -f python scripts/tile_one.py --path auto --c $1 --h $2 --n 1024 --bwd --reps 1 > $OUT/ncu_block_$1.log 2>&1
  tail -1 $OUT/ncu_block_$1.log
done
# launch list of the bench command (hot-path legs only): per-launch gpu time, cold cache, serialised
python bench.py --steps 2 --warmup 3 --no-train --no-stats > $OUT/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r2_ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-train --no-stats > $OUT/ncu_bench.log 2>&1
tail -2 $OUT/ncu_bench.log
