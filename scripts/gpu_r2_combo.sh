#!/bin/bash
# one call on a 2-GPU box: (1) single-GPU parity with the pipelined merge kernel + A/B benches, (2) 2-GPU benches
mkdir -p gpurun_out
echo "== pytest -m gpu with DPK_OPTIONS=agg_pipe=1"
DPK_OPTIONS=agg_pipe=1 timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_pipe.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_pipe.log
echo "== pytest kernels default options"
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_subbuckets.py tests/test_gpu_rdd.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
run1() { # tag, args...
  tag=$1; shift
  echo "== bench1 $tag: $@"
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-parity "$@" > gpurun_out/bench1_$tag.json 2> gpurun_out/bench1_$tag.err; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench1_$tag.json"))
    print("ms/step", round(d["ms_per_step"],3), "value %.3e"%d["value"])
    for k in d["kernels"][:6]: print("  ", k["kernel"], k["ms"], "alg_gbs", k["alg_gbs"])
except Exception as e:
    print("no json:", e); print(open("gpurun_out/bench1_$tag.err").read()[-2000:])
PY
}
run1 base
run1 pipe1 --agg-pipe 1
run1 pipe2 --agg-pipe 2
run1 pipe1c4 --agg-pipe 1 --config c4 --rows-per-gpu 200000000
run1 basec4 --config c4 --rows-per-gpu 200000000
bash scripts/gpu_r2_multi.sh 2 overlap c4small spmd
