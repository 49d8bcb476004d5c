#!/bin/bash
mkdir -p gpurun_out
echo "== pytest kernels"
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_subbuckets.py tests/test_gpu_host_api.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
for cfg in "512 3" "256 3" "512 4"; do
  set -- $cfg
  echo "== bench scatter_threads=$1 agg_ctas=$2"
  timeout 600 python scripts/bench_ab.py --steps 10 --warmup 3 --e2e-steps 1 --e2e-depth 1 --no-cpu-baseline --scatter-threads $1 --agg-ctas $2 > gpurun_out/bench_t$1_c$2.json 2> gpurun_out/bench_t$1_c$2.err; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_t$1_c$2.json"))
    print("ms/step", round(d["ms_per_step"],3), "value %.3e"%d["value"])
    for k in d["kernels"][:6]: print("  ", k["kernel"], round(k["ms_per_step"],3), "alg_gbs", k["alg_gbs"] and round(k["alg_gbs"]))
except Exception as e:
    print("no json:", e); print(open("gpurun_out/bench_t$1_c$2.err").read()[-2000:])
PY
done
