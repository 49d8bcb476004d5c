#!/bin/bash
# One --set full capture of one kernel: $1 = kernel regex, $2 = tag, $3 = launches to skip, rest = bench args
# (COUNT=n in the environment captures n consecutive launches)
K=$1; TAG=$2; SKIP=$3; shift 3
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c ${COUNT:-1} \
    -f -o gpurun_out/${TAG} python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline "$@" \
    > gpurun_out/${TAG}.log 2>&1
echo "$K rc=$?"; ls -la gpurun_out/${TAG}.ncu-rep 2>/dev/null
