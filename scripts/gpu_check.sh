#!/bin/bash
# Runs on the GPU box (via gpurun): smoke, GPU parity tests, bench, ncu launch list.
# Everything is logged under gpurun_out/; a failing stage does not stop the later ones.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
tail -3 gpurun_out/smoke.log
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${RUN_NCU:-1}" = "1" ]; then
  echo "== ncu launch list"
  timeout 300 python bench.py --no-train --steps 2 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/plain.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
      python bench.py --no-train --steps 2 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/ncu.log 2>&1
  echo "ncu exit $?"; tail -3 gpurun_out/ncu.log
fi
