#!/bin/bash
# Multi-GPU scaling runs on one box.  usage: gpu_scale.sh "N exchange" ["N exchange" ...]
mkdir -p gpurun_out
port=29520
for cfg in "$@"; do
  set -- $cfg
  N=$1; EX=$2; port=$((port+1))
  echo "== bench --gpus $N --exchange $EX"
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --e2e-steps 2 --no-cpu-baseline \
      > gpurun_out/scale_${N}_${EX}.json 2> gpurun_out/scale_${N}_${EX}.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $port bench.py --gpus $N --steps 5 --warmup 3 --e2e-steps 2 --exchange $EX \
      > gpurun_out/scale_${N}_${EX}.json 2> gpurun_out/scale_${N}_${EX}.err
  fi
  echo "rc=$?"; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/scale_${N}_${EX}.err | tail -3 | cut -c1-300
  python - <<EOF
import json
try:
    b = json.load(open("gpurun_out/scale_${N}_${EX}.json"))
    print("  value %.3e rows/s  %.2f ms/step  e2e %.3e" % (b["value"], b["ms_per_step"], b["e2e"]["value"]))
    if b.get("roofline_exchange"):
        r = b["roofline_exchange"]; print("  exchange: %.0f GB/s/GPU/dir, frac %.2f, %.2f ms/step" % (r["achieved"], r["frac"], r["ms_per_step"]))
    for k in b["kernels"][:6]:
        print("   %-14s %7.3f ms" % (k["kernel"], k["ms_per_step"]))
except Exception as e:
    print("  failed:", e)
EOF
done
