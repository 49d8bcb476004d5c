#!/bin/bash
# Final ncu passes of a round (1 GPU).  $1 = tag.  Keeps gpurun_out/ under the 64 MiB return limit:
# raw-page CSVs are exported on the box, the .ncu-rep files come back only while they fit.
TAG=${1:-r01f}
rm -rf gpurun_out; mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --e2e-steps 1 --e2e-depth 1 --no-cpu-baseline"
# (1) launch list: device time per launch (cold-cache, serialised: compare SHARES with bench.py's events)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_launches_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/${TAG}_launches.csv
# (2) --set full: the aggregate kernel (1 launch per step), the last map-side scatter of a step and the
#     segmented scatter that follows it, the matching histogram launches
cap() {  # regex tag skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s $3 -c $4 \
      -f -o gpurun_out/${TAG}_$2 $BENCH > gpurun_out/${TAG}_$2.log 2>&1
  echo "$1 rc=$?"
  ncu -i gpurun_out/${TAG}_$2.ncu-rep --page raw --csv > gpurun_out/${TAG}_$2_raw.csv 2>/dev/null
  ls -la gpurun_out/${TAG}_$2.ncu-rep
}
cap k_smem_aggregate agg 3 1
cap k_part_scatter scatter 34 2
cap k_part_count count 34 2
cap k_copy_segments copy 0 1 2>/dev/null
# size guard
while [ $(du -sm gpurun_out | cut -f1) -ge 60 ]; do
  big=$(ls -S gpurun_out/*.ncu-rep 2>/dev/null | head -1); [ -z "$big" ] && break
  echo "dropping $big to stay under the return limit"; rm -f "$big"
done
du -sh gpurun_out
