#!/bin/bash
# round 2: parity tests + A/B of the new multisplit / merge kernels (one GPU)
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  echo "== bench scatter_bulk=$1 agg_impl=$2"
  timeout 600 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --e2e-depth 1 --no-cpu-baseline --scatter-bulk $1 --agg-impl $2 > gpurun_out/bench_sb$1_ag$2.json 2> gpurun_out/bench_sb$1_ag$2.err; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_sb$1_ag$2.json"))
    print("ms/step", round(d["ms_per_step"],3), "value %.3e"%d["value"])
    for k in d["kernels"][:8]: print("  ", k["kernel"], round(k["ms_per_step"],3), "alg_gbs", k["alg_gbs"] and round(k["alg_gbs"]))
except Exception as e:
    print("no json:", e); print(open("gpurun_out/bench_sb$1_ag$2.err").read()[-2000:])
PY
done
