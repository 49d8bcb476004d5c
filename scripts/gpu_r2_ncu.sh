#!/bin/bash
# round 2 ncu captures (1 GPU): $1 = tag; launch list + --set full of the merge kernel and the bulk multisplit
TAG=${1:-r02f}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_launches_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/${TAG}_launches.csv
cap() {  # regex tag skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s $3 -c $4 \
      -f -o gpurun_out/${TAG}_$2 $BENCH > gpurun_out/${TAG}_$2.log 2>&1
  echo "$1 rc=$?"
  ncu -i gpurun_out/${TAG}_$2.ncu-rep --page raw --csv > gpurun_out/${TAG}_$2_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_$2.ncu-rep --page source --csv > gpurun_out/${TAG}_$2_source.csv 2>/dev/null
  ls -la gpurun_out/${TAG}_$2.ncu-rep
  rm -f gpurun_out/${TAG}_$2.ncu-rep
}
cap k_smem_aggregate agg 6 2
cap k_part_scatter scatter 6 2
cap k_part_count count 6 2
du -sh gpurun_out
