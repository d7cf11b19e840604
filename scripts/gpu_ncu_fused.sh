#!/bin/bash
# GPU box: one ncu --set full capture of the fused forward/backward kernels (128x28^2, N=256).
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
CMD="python scripts/sweep.py --batches ${NCU_BATCH:-256} --iters 1 --variants ${NCU_VARIANT:-fused_cs4_t512} --out gpurun_out/sweep_ncu.json"
timeout 300 $CMD > gpurun_out/ncu_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:${NCU_KERNEL:-fused} -s ${NCU_SKIP:-4} -c ${NCU_COUNT:-4} -f -o gpurun_out/${NCU_OUT:-fused_prof} $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
