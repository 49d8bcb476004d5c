#!/bin/bash
# last single-GPU run of the round: A/B of the split merge launch, parity suite, bench lines, ncu captures
mkdir -p gpurun_out
run1() { # tag, args...
  tag=$1; shift
  echo "== bench1 $tag: $@ (DPK_OPTIONS=$DPK_OPTIONS)"
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
except Exception as e:
    print("no json:", e); print(open("gpurun_out/bench1_$tag.err").read()[-2500:])
PY
}
Q="--no-e2e --no-cpu-baseline --no-parity"
run1 ab_split1 $Q
DPK_OPTIONS=agg_split=0 run1 ab_split0 $Q
DPK_OPTIONS=agg_ctas=3 run1 ab_split1_ctas3 $Q
run1 ab_c4_split1 --config c4 --steps 5 $Q
DPK_OPTIONS=agg_split=0 run1 ab_c4_split0 --config c4 --steps 5 $Q
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
run1 c2_final
run1 c4_final --config c4 --steps 5 --no-cpu-baseline --e2e-steps 1 --e2e-depth 2
run1 c3_final --config c3 --steps 5 --no-cpu-baseline --no-e2e
bash scripts/gpu_r2_ncu.sh r02g
