#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log
run() { # tag, args...
  tag=$1; shift
  echo "== bench $tag: $@"
  timeout 900 python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$tag.json"))
    print("ms/step", round(d["ms_per_step"],3), "value %.3e"%d["value"], "parity", d.get("parity",{}).get("parity_checked"), d.get("parity",{}).get("seconds"))
    for k in d["kernels"][:6]: print("  ", k["kernel"], k["ms"], "alg_gbs", k["alg_gbs"])
    if d.get("e2e"): print("  e2e ms/step", round(d["e2e"]["ms_per_step"],2), "value %.3e"%d["e2e"]["value"])
    if d.get("cpu_baseline"): print("  cpu", d["cpu_baseline"]["kind"], "%.3e"%d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
except Exception as e:
    print("no json:", e); print(open("gpurun_out/bench_$tag.err").read()[-3000:])
PY
}
run full
run cur0 --agg-cursor 0 --no-e2e --no-cpu-baseline --no-parity
run cur1c4 --agg-ctas 4 --no-e2e --no-cpu-baseline --no-parity
run sb4 --sub-bits 4 --no-e2e --no-cpu-baseline --no-parity
run sb6 --sub-bits 6 --no-e2e --no-cpu-baseline --no-parity
echo "== reference arm"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; head -c 1500 gpurun_out/bench_ref.json
