#!/bin/bash
# final single-GPU state of the round: parity, the pipelined merge kernel under test options, bench lines, ncu captures
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
echo "== pytest kernels with agg_pipe=1 (probe-loop insert)"
DPK_OPTIONS=agg_pipe=1 timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_subbuckets.py tests/test_gpu_host_api.py -m gpu -x -q > gpurun_out/pytest_gpu_pipe.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_pipe.log
run1() { # tag, args...
  tag=$1; shift
  echo "== bench1 $tag: $@"
  timeout 900 python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/bench1_$tag.json 2> gpurun_out/bench1_$tag.err; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench1_$tag.json"))
    print("ms/step", round(d["ms_per_step"],3), "value %.3e"%d["value"], "parity", d.get("parity",{}).get("parity_checked"))
    for k in d["kernels"][:6]: print("  ", k["kernel"], k["n"], k["ms"], "alg_gbs", k["alg_gbs"])
    for r in ("roofline","roofline_map_scatter","roofline_map_side","roofline_reduce"):
        if d.get(r): print("  ", r, round(d[r]["frac"],3))
    if d.get("e2e"): print("   e2e ms/step", round(d["e2e"]["ms_per_step"],2), "value %.3e"%d["e2e"]["value"])
except Exception as e:
    print("no json:", e); print(open("gpurun_out/bench1_$tag.err").read()[-2500:])
PY
}
run1 c2_final
run1 c2_pipe --no-e2e --no-cpu-baseline --no-parity --agg-pipe 1
run1 c4_final --config c4 --steps 5 --no-cpu-baseline --e2e-steps 1 --e2e-depth 2
run1 c4_pipe --config c4 --steps 5 --no-cpu-baseline --no-e2e --no-parity --agg-pipe 1
run1 c3_final --config c3 --steps 5 --no-cpu-baseline --no-e2e
bash scripts/gpu_r2_ncu.sh r02f
