#!/bin/bash
# ncu passes on the GPU box (1 GPU).  $1 = tag for output names.
# (1) launch list with device time per launch (cold-cache, serialised: compare SHARES)
# (2) --set full captures of the dominant kernels
TAG=${1:-r01}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_launches_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/${TAG}_launches.csv
for K in k_part_scatter k_tbl_insert k_part_count k_tbl_compact; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 10 -c 2 \
      -f -o gpurun_out/${TAG}_$K $BENCH > gpurun_out/${TAG}_${K}.log 2>&1
  echo "$K rc=$?"; ls -la gpurun_out/${TAG}_$K.ncu-rep 2>/dev/null
done
