#!/bin/bash
# multi-GPU checks: $1 = number of GPUs.  Parity script + bench lines (c2, c4, c3) at N GPUs.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
echo "== multi_gpu_check N=$N"
timeout 900 $TR scripts/multi_gpu_check.py > gpurun_out/multi_gpu_check_n$N.log 2>&1; echo "rc=$?"
grep -E "^case|MISMATCH|Error|error" gpurun_out/multi_gpu_check_n$N.log | head -20
run() { # tag, args...
  tag=$1; shift
  echo "== bench $tag: $@"
  timeout 1200 $TR bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${tag}_n$N.json"))
    print("ms/step", round(d["ms_per_step"],3), "value %.3e"%d["value"], "parity", d.get("parity",{}).get("parity_checked"), d.get("parity",{}).get("seconds"))
    for k in d["kernels"][:8]: print("  ", k["kernel"], k["n"], k["ms"], "alg_gbs", k["alg_gbs"])
    if d.get("roofline_exchange"): print("  exchange", round(d["roofline_exchange"]["achieved"]), "GB/s", round(d["roofline_exchange"]["ms_per_step_max_over_ranks"],3), "ms")
    if d.get("e2e"): print("  e2e ms/step", round(d["e2e"]["ms_per_step"],2), "value %.3e"%d["e2e"]["value"])
except Exception as e:
    print("no json:", e); print(open("gpurun_out/bench_${tag}_n$N.err").read()[-3000:])
PY
}
run c2
shift
for extra in "$@"; do
  case $extra in
    c4) run c4 --config c4 --steps 5 --e2e-steps 1 --e2e-depth 2 ;;
    c4mc) run c4mc --config c4 --steps 5 --map-combine --no-e2e ;;
    c4small) run c4small --config c4 --rows-per-gpu 100000000 --e2e-steps 1 ;;
    c3) run c3 --config c3 --steps 3 --e2e-steps 1 --e2e-depth 2 ;;
    c3small) run c3small --config c3 --rows-per-gpu 20000000 --steps 3 --e2e-steps 1 ;;
    nccl) run c2nccl --exchange nccl --no-e2e ;;
    pcie) echo "== pcie probe, all ranks at once"; timeout 300 $TR scripts/pcie_probe_multi.py 2>&1 | grep -E "alone|both" ;;
    overlap) run c2ov2 --overlap-push 2 --no-e2e ;;
    spmd) echo "== spmd_check"; timeout 900 $TR scripts/spmd_check.py > gpurun_out/spmd_check_n$N.log 2>&1; echo "rc=$?"; grep -E "^wc|^pagerank|Error|error" gpurun_out/spmd_check_n$N.log | head ;;
  esac
done
