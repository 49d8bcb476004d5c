#!/bin/bash
# quick kernel-only sweeps (no e2e leg worth mentioning): each argument = extra bench.py flags
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --e2e-steps 1 --e2e-depth 1 --no-cpu-baseline"
i=0
for extra in "$@"; do
  i=$((i+1))
  timeout 300 $B $extra > gpurun_out/q_$i.json 2> gpurun_out/q_$i.err
  python - <<EOF2
import json
try:
    b = json.load(open("gpurun_out/q_$i.json"))
    ks = {k["kernel"]: k["ms_per_step"] for k in b["kernels"]}
    print("%-28s %.2f ms/step  agg %.3f  scat %.3f  seg %.3f  cnt %.3f  segcnt %.3f" % ("$extra", b["ms_per_step"],
          ks.get("smem_aggregate", 0), ks.get("part_scatter", 0), ks.get("seg_scatter", 0), ks.get("part_count", 0), ks.get("seg_count", 0)))
except Exception as e:
    print("$extra failed:", e); print(open("gpurun_out/q_$i.err").read()[-800:])
EOF2
done
