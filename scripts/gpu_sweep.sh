#!/bin/bash
# GPU box: parity tests first (fail fast), then the kernel-path sweep.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -8 gpurun_out/pytest_gpu.log
echo "== sweep"; timeout 1200 python scripts/sweep.py ${SWEEP_ARGS:-} > gpurun_out/sweep.log 2>&1; echo "sweep exit $?"
cat gpurun_out/sweep.log | tail -80
