#!/bin/bash
# session 2: exchange forms at N GPUs ($1): parity script, then C2 with push / fused (TMA bulk stores into peer memory) /
# overlapped push; extra tags select more lines
N=${1:-2}
shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
echo "== new single-GPU tests"
timeout 600 python -m pytest tests/test_gpu_subbuckets.py -m gpu -x -q -k "pointer_mode or fused_plan" 2>&1 | tail -3
echo "== multi_gpu_check N=$N"
timeout 900 $TR scripts/multi_gpu_check.py > gpurun_out/multi_gpu_check_n$N.log 2>&1; echo "rc=$?"
grep -E "^case|MISMATCH|Error|error" gpurun_out/multi_gpu_check_n$N.log | head -20
run() { # tag, args...
  tag=$1; shift
  echo "== bench $tag: $@ (DPK_OPTIONS=$DPK_OPTIONS)"
  timeout 1200 $TR bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err; echo "rc=$?"
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
Q="--no-e2e"
for extra in "$@"; do
  case $extra in
    push) run c2push --exchange push --pipeline off $Q ;;
    p1x2b) run c2p1x2b --pipeline 1x2 --copy-engine 2 $Q ;;
    p2x2b) run c2p2x2b --pipeline 2x2 --copy-engine 2 $Q ;;
    p2x1b) run c2p2x1b --pipeline 2x1 --copy-engine 2 $Q ;;
    p1x1b) run c2p1x1b --pipeline 1x1 --copy-engine 2 $Q ;;
    fused) run c2fused --exchange fused --pipeline off $Q ;;
    fused512) DPK_OPTIONS=scatter_ptr_threads=512 run c2fused512 --exchange fused $Q ;;
    fusedold) DPK_OPTIONS=scatter_ptr_bulk=0 run c2fusedold --exchange fused $Q ;;
    ov2) run c2ov2 --exchange push --overlap-push 2 $Q ;;
    ov4) run c2ov4 --exchange push --overlap-push 4 $Q ;;
    ov2s0) run c2ov2s0 --exchange push --overlap-push 2 --copy-sms 0 $Q ;;
    ov2s16) run c2ov2s16 --exchange push --overlap-push 2 --copy-sms 16 $Q ;;
    ov4s16) run c2ov4s16 --exchange push --overlap-push 4 --copy-sms 16 $Q ;;
    ov4s32) run c2ov4s32 --exchange push --overlap-push 4 --copy-sms 32 $Q ;;
    pushs16) DPK_OPTIONS=copy_sms=16 run c2pushs16 --exchange push $Q ;;
    pushs32) DPK_OPTIONS=copy_sms=32 run c2pushs32 --exchange push $Q ;;
    pushsb4) run c2pushsb4 --exchange push --sub-bits 4 $Q ;;
    fusedsb4) run c2fusedsb4 --exchange fused --sub-bits 4 --pipeline off $Q ;;
    text) echo "== text ingest tests + wc_e2e"; timeout 600 python -m pytest tests/test_gpu_textingest.py tests/test_gpu_rdd.py -m gpu -x -q 2>&1 | tail -3
          timeout 300 python scripts/wc_e2e.py 2>&1 | tail -3; timeout 300 python scripts/wc_e2e.py 1000000 rowwise 2>&1 | tail -3 ;;
    p1x2) run c2p1x2 --pipeline 1x2 --copy-engine 0 $Q ;;
    p2x2) run c2p2x2 --pipeline 2x2 $Q ;;
    p4x2) run c2p4x2 --pipeline 4x2 $Q ;;
    p4x1) run c2p4x1 --pipeline 4x1 $Q ;;
    p4x4) run c2p4x4 --pipeline 4x4 $Q ;;
    p4x2s32) run c2p4x2s32 --pipeline 4x2 --copy-sms 32 $Q ;;
    p4x2s12) run c2p4x2s12 --pipeline 4x2 --copy-sms 12 $Q ;;
    p1x1ce) run c2p1x1ce --pipeline 1x1 $Q ;;
    p1x2ce) run c2p1x2ce --pipeline 1x2 $Q ;;
    p2x2ce) run c2p2x2ce --pipeline 2x2 $Q ;;
    p2x4ce) run c2p2x4ce --pipeline 2x4 $Q ;;
    p4x2ce) run c2p4x2ce --pipeline 4x2 $Q ;;
    p2x2sm32) run c2p2x2sm32 --pipeline 2x2 --copy-engine 0 --copy-sms 32 $Q ;;
    c2) run c2 ;;
    c4) run c4 --config c4 --steps 5 --e2e-steps 1 --e2e-depth 2 ;;
    c4fused) run c4fused --config c4 --steps 5 --exchange fused $Q ;;
    c4small) run c4small --config c4 --rows-per-gpu 100000000 --e2e-steps 1 ;;
    c3) run c3 --config c3 --steps 3 --e2e-steps 1 --e2e-depth 2 ;;
    c3small) run c3small --config c3 --rows-per-gpu 20000000 --steps 3 --e2e-steps 1 ;;
    nccl) run c2nccl --exchange nccl $Q ;;
    spmd) echo "== spmd_check"; timeout 900 $TR scripts/spmd_check.py > gpurun_out/spmd_check_n$N.log 2>&1; echo "rc=$?"; grep -E "^wc|^pagerank|Error|error" gpurun_out/spmd_check_n$N.log | head ;;
  esac
done
