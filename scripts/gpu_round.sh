#!/bin/bash
# one GPU call: GPU parity tests, then bench variants (each argument = extra bench.py flags)
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu.log
bash scripts/gpu_sweep.sh "$@"
