#!/bin/bash
# session 2, last single-GPU run: the whole parity suite, the default bench line, C4 / C3 lines, C1 (wc, device tokeniser
# vs row-wise) and C5-shape (pagerank) end to end, launch list + one --set full capture of the merge kernel
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
run1() { # tag, args...
  tag=$1; shift
  echo "== bench1 $tag: $@"
  timeout 900 python bench.py "$@" > gpurun_out/bench1_$tag.json 2> gpurun_out/bench1_$tag.err; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench1_$tag.json"))
    print("ms/step", round(d["ms_per_step"],3), "value %.3e"%d["value"], "parity", d.get("parity",{}).get("parity_checked"))
    for k in d["kernels"][:7]: print("  ", k["kernel"], k["n"], k["ms"], "alg_gbs", k["alg_gbs"])
    for r in ("roofline","roofline_map_scatter","roofline_map_side","roofline_reduce"):
        if d.get(r): print("  ", r, round(d[r]["frac"],3))
    if d.get("e2e"): print("   e2e ms/step", round(d["e2e"]["ms_per_step"],2), "value %.3e"%d["e2e"]["value"])
    if d.get("cpu_baseline"): print("   cpu", d["cpu_baseline"].get("value"), d.get("clocks"))
except Exception as e:
    print("no json:", e); print(open("gpurun_out/bench1_$tag.err").read()[-2500:])
PY
}
run1 c2_final
run1 c4_final --config c4 --steps 5 --no-cpu-baseline --e2e-steps 1 --e2e-depth 2
run1 c3_final --config c3 --steps 5 --no-cpu-baseline --no-e2e
echo "== wc_e2e (C1): device tokeniser, then row-wise"
timeout 300 python scripts/wc_e2e.py > gpurun_out/wc_e2e_device.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/wc_e2e_device.log
timeout 300 python scripts/wc_e2e.py 1000000 rowwise > gpurun_out/wc_e2e_rowwise.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/wc_e2e_rowwise.log
echo "== pagerank_e2e (C5 shape, 3e5 vertices / 3e6 edges)"
timeout 400 python scripts/pagerank_e2e.py 300000 3000000 > gpurun_out/pagerank_e2e.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pagerank_e2e.log
echo "== ncu: launch list + merge kernel"
TAG=r02g
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_launches_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/${TAG}_launches.csv
cap() {  # regex tag skip count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s $3 -c $4 \
      -f -o gpurun_out/${TAG}_$2 $BENCH > gpurun_out/${TAG}_$2.log 2>&1
  echo "$1 rc=$?"
  ncu -i gpurun_out/${TAG}_$2.ncu-rep --page raw --csv > gpurun_out/${TAG}_$2_raw.csv 2>/dev/null
  rm -f gpurun_out/${TAG}_$2.ncu-rep
}
cap k_smem_aggregate agg 6 1
cap k_part_scatter scatter 6 2
du -sh gpurun_out
