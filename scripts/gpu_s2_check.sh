#!/bin/bash
# session-2 single-GPU check of HEAD: parity suite, the default bench line, C1 (wc) and C5 (pagerank) end to end
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
echo "== bench c2 (default flags)"
timeout 900 python bench.py > gpurun_out/bench_c2_n1.json 2> gpurun_out/bench_c2_n1.err; echo "rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_c2_n1.json"))
    print("ms/step", round(d["ms_per_step"],3), "value %.3e"%d["value"], "parity", d.get("parity",{}).get("parity_checked"))
    for k in d["kernels"][:7]: print("  ", k["kernel"], k["n"], k["ms"], "alg_gbs", k["alg_gbs"])
    for r in ("roofline","roofline_map_scatter","roofline_map_side","roofline_reduce"):
        if d.get(r): print("  ", r, round(d[r]["frac"],3))
    if d.get("e2e"): print("   e2e ms/step", round(d["e2e"]["ms_per_step"],2), "value %.3e"%d["e2e"]["value"])
    print("   cpu", d.get("cpu_baseline",{}).get("value"), d.get("clocks"))
except Exception as e:
    print("no json:", e); print(open("gpurun_out/bench_c2_n1.err").read()[-2500:])
PY
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; head -c 1200 gpurun_out/bench_ref.json
echo "== wc_e2e (C1)"
timeout 600 python scripts/wc_e2e.py > gpurun_out/wc_e2e.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/wc_e2e.log
echo "== pagerank_e2e (C5 shape, 1e5 vertices / 1e6 edges)"
timeout 600 python scripts/pagerank_e2e.py > gpurun_out/pagerank_e2e.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pagerank_e2e.log
