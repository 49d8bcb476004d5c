#!/bin/bash
# round 2 ncu captures (1 GPU): $1 = tag; captures the merge kernel and the bulk multisplit (map + segmented)
TAG=${1:-r02a}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --e2e-steps 1 --e2e-depth 1 --no-cpu-baseline"
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
while [ $(du -sm gpurun_out | cut -f1) -ge 60 ]; do
  big=$(ls -S gpurun_out/*.ncu-rep 2>/dev/null | head -1); [ -z "$big" ] && break
  echo "dropping $big to stay under the return limit"; rm -f "$big"
done
du -sh gpurun_out
