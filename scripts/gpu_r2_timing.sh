#!/bin/bash
mkdir -p gpurun_out
for opt in "agg_timing=1" "agg_timing=1,agg_batched=0" "agg_timing=1,agg_batched=0,agg_ctas=4"; do
  echo "== $opt"
  DPK_OPTIONS=$opt timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity 2>&1 >/dev/null | grep agg_timing | tail -2
done
echo "== c4 200M rows"
DPK_OPTIONS=agg_timing=1 timeout 600 python bench.py --config c4 --rows-per-gpu 200000000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity 2>&1 >/dev/null | grep agg_timing | tail -2
