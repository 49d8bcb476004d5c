#!/bin/bash
# single-GPU validation of the round's final state: full parity suite, bench lines (c2 full, c4, c3), timing, A/Bs
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
run1() { # tag, args...
  tag=$1; shift
  echo "== bench1 $tag: $@"
  timeout 900 python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/bench1_$tag.json 2> gpurun_out/bench1_$tag.err; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench1_$tag.json"))
    print("ms/step", round(d["ms_per_step"],3), "value %.3e"%d["value"], "parity", d.get("parity",{}).get("parity_checked"))
    for k in d["kernels"][:7]: print("  ", k["kernel"], k["n"], k["ms"], "alg_gbs", k["alg_gbs"])
    for r in ("roofline","roofline_map_scatter","roofline_map_side","roofline_reduce"):
        if d.get(r): print("  ", r, round(d[r]["frac"],3))
    if d.get("e2e"): print("   e2e ms/step", round(d["e2e"]["ms_per_step"],2), "value %.3e"%d["e2e"]["value"])
    if d.get("cpu_baseline"): print("   cpu", d["cpu_baseline"]["kind"], "%.3e"%d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
except Exception as e:
    print("no json:", e); print(open("gpurun_out/bench1_$tag.err").read()[-2500:])
PY
}
run1 c2_full
run1 c2_persplit --no-e2e --no-cpu-baseline --no-parity --map-splits 1
run1 c4 --config c4 --steps 5 --no-cpu-baseline --e2e-steps 1 --e2e-depth 2
run1 c4_pipe --config c4 --steps 5 --no-cpu-baseline --no-e2e --no-parity --agg-pipe 1
run1 c4_mc --config c4 --steps 5 --no-cpu-baseline --no-e2e --map-combine
run1 c3 --config c3 --steps 5 --no-cpu-baseline --e2e-steps 1 --e2e-depth 2
echo "== agg timing"
bash scripts/gpu_r2_timing.sh
echo "== wc e2e / pagerank e2e"
timeout 600 python scripts/wc_e2e.py 2>&1 | tail -3
timeout 900 python scripts/pagerank_e2e.py 2>&1 | tail -3
