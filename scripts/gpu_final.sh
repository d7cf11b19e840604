#!/bin/bash
# Round-end evidence run on one B200: bench line, sweep table, ncu launch list, one full capture.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref exit $?"
echo "== sweep"; timeout 600 python scripts/sweep.py --batches 32,64,128,256,512,1024 --variants auto,stream_nochunk > gpurun_out/sweep.log 2>&1; echo "sweep exit $?"
echo "== ncu launch list"
CMD="python bench.py --no-train --steps 2 --warmup 3"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu list exit $?"
echo "== ncu full (dominant kernel)"
CMD2="python scripts/sweep.py --batches 1024 --iters 1 --variants auto --out gpurun_out/sweep_ncu.json"
timeout 300 $CMD2 > gpurun_out/ncu_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:l2_bwd -c 2 -f -o gpurun_out/dominant_prof $CMD2 > gpurun_out/ncu_full2.log 2>&1
echo "ncu full exit $?"
echo "== ncu full (tcgen05 FC GEMM)"
CMD3="python scripts/gemm_accuracy.py"
timeout 300 $CMD3 > gpurun_out/gemm_accuracy.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:gemm_umma -c 2 -f -o gpurun_out/umma_gemm_prof $CMD3 > gpurun_out/ncu_full3.log 2>&1
echo "ncu umma exit $?"
