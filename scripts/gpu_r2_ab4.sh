#!/bin/bash
mkdir -p gpurun_out
echo "== pytest (kernels, host api, rdd) with scatter_threads=1024,agg_batched=0"
DPK_OPTIONS=scatter_threads=1024,agg_batched=0 timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_subbuckets.py tests/test_gpu_host_api.py tests/test_gpu_rdd.py -m gpu -x -q > gpurun_out/pytest_gpu_t1024.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_t1024.log
run1() { # tag, args...
  tag=$1; shift
  echo "== bench1 $tag: $@"
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-parity "$@" > gpurun_out/bench1_$tag.json 2> gpurun_out/bench1_$tag.err; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench1_$tag.json"))
    print("ms/step", round(d["ms_per_step"],3), "value %.3e"%d["value"])
    for k in d["kernels"][:5]: print("  ", k["kernel"], k["ms"], "alg_gbs", k["alg_gbs"])
except Exception as e:
    print("no json:", e); print(open("gpurun_out/bench1_$tag.err").read()[-2000:])
PY
}
run1 base
run1 nobatch --agg-batched 0
run1 nobatch_c4 --agg-batched 0 --agg-ctas 4
run1 t1024 --scatter-threads 1024
run1 t1024_sb4 --scatter-threads 1024 --sub-bits 4
run1 t1024_sb6 --scatter-threads 1024 --sub-bits 6
