#!/bin/bash
# bench sweeps: each argument is a quoted string of extra bench.py flags
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
i=0
for extra in "$@"; do
  i=$((i+1))
  echo "== bench $extra"
  timeout 600 $B $extra > gpurun_out/sweep_$i.json 2> gpurun_out/sweep_$i.err; echo "rc=$?"
  python - <<EOF
import json
try:
    b = json.load(open("gpurun_out/sweep_$i.json"))
    ms = b.get("roofline_map_scatter") or b["roofline"]
    print("  value %.3e rows/s  %.2f ms/step  e2e %.3e  map_scatter_frac %.3f  dominant %s frac %.3f" % (
        b["value"], b["ms_per_step"], b["e2e"]["value"], ms["frac"], b["roofline"]["kernel"][:18], b["roofline"]["frac"]))
    for k in b["kernels"][:6]:
        print("   %-14s %7.3f ms  share %.3f" % (k["kernel"], k["ms_per_step"], k["share"]))
except Exception as e:
    print("  failed:", e); print(open("gpurun_out/sweep_$i.err").read()[-1500:])
EOF
done
