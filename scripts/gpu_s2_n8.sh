#!/bin/bash
# session 2, the 8-GPU evidence run ($1 = GPUs, default 8): what the driver's own scaling run does not cover -- the parity
# script at N ranks, C4 and C3 lines, the operator surface (wc + Bagel PageRank) with one driver per GPU -- plus one short
# C2 line in the default mode (that the default works at N before the driver runs it).
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
echo "== multi_gpu_check N=$N"
timeout 600 $TR scripts/multi_gpu_check.py > gpurun_out/multi_gpu_check_n$N.log 2>&1; echo "rc=$?"
grep -E "^case|MISMATCH|Error|error" gpurun_out/multi_gpu_check_n$N.log | head -20
run() { # tag, args...
  tag=$1; shift
  echo "== bench $tag: $@"
  timeout 600 $TR bench.py --gpus $N "$@" > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${tag}_n$N.json"))
    print("ms/step", round(d["ms_per_step"],3), "value %.3e"%d["value"], "parity", d.get("parity",{}).get("parity_checked"), d.get("parity",{}).get("seconds"))
    for k in d["kernels"][:8]: print("  ", k["kernel"], k["n"], k["ms"], "alg_gbs", k["alg_gbs"])
    rx = d.get("roofline_exchange") or {}
    if rx.get("achieved"): print("  exchange", round(rx["achieved"]), "GB/s", round(rx["ms_per_step_max_over_ranks"],3), "ms")
    elif rx: print("  exchange by copy engines: step - kernels =", round(rx["step_ms_minus_kernel_ms"],3), "ms; NVLink lower bound", round(rx["lower_bound_ms"],3), "ms")
    if d.get("e2e"): print("  e2e ms/step", round(d["e2e"]["ms_per_step"],2), "value %.3e"%d["e2e"]["value"])
except Exception as e:
    print("no json:", e); print(open("gpurun_out/bench_${tag}_n$N.err").read()[-3000:])
PY
}
run c2 --steps 10 --warmup 3 --no-e2e
run c3 --config c3 --steps 3 --warmup 3 --no-e2e
run c4 --config c4 --steps 3 --warmup 3 --no-e2e
