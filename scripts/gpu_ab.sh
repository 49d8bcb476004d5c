#!/bin/bash
# A/B of reduce-side implementations and sub-bucket counts on one GPU box.
# usage: gpu_ab.sh ["impl subbits" ...]      (default list below)
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
B="python bench.py --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
if [ $# -eq 0 ]; then set -- "2 -1" "2 4" "2 6" "2 3" "1 6"; fi
for cfg in "$@"; do
  set -- $cfg
  echo "== bench reduce_impl=$1 sub_bits=$2"
  timeout 600 $B --reduce-impl $1 --sub-bits $2 > gpurun_out/ab_$1_$2.json 2> gpurun_out/ab_$1_$2.err; echo "rc=$?"
  python - <<EOF
import json
try:
    b = json.load(open("gpurun_out/ab_$1_$2.json"))
    print("  value %.3e rows/s  %.2f ms/step  e2e %.3e" % (b["value"], b["ms_per_step"], b["e2e"]["value"]))
    for k in b["kernels"]:
        print("   %-14s %7.3f ms  share %.3f" % (k["kernel"], k["ms_per_step"], k["share"]))
except Exception as e:
    print("  failed:", e); print(open("gpurun_out/ab_$1_$2.err").read()[-1500:])
EOF
done
