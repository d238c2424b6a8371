#!/bin/bash
# last visit: the whole GPU test-suite, smoke, bench (own arm, bf16), refreshed layer timing and launch list
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; grep -n "passed\|failed" gpurun_out/pytest_gpu.txt | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.txt
( time timeout 900 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
( timeout 600 python bench.py --dtype bf16 --no-cpu-baseline ) > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bf16 rc=$?"
timeout 300 python tools/e2e_breakdown.py > gpurun_out/e2e_breakdown.txt 2>&1; echo "e2e breakdown rc=$?"
export PDES_EXEC_GRAPH=0
PDES_WGRAD_STREAMS=0 timeout 300 python tools/layer_timing.py > gpurun_out/layer_timing.out 2> gpurun_out/layer_timing_s0.txt; echo "timing rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py --steps 3 > gpurun_out/prof_step.log 2>&1; echo "list rc=$?"
python tools/launch_summary.py gpurun_out/launches.csv 3 30 > gpurun_out/launches.txt 2>&1
unset PDES_EXEC_GRAPH
python - <<'PY'
import json
for f in ("bench", "bench_bf16"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, {k: d[k] for k in ("value", "ms_per_step", "launches_per_step", "dtype")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("stock_adam"),
              "roof", d["roofline"]["frac"], d["roofline_step"]["frac"], "stencil", d["roofline_stencil"]["frac"], d.get("gpu_library_baseline"), d.get("cpu_baseline", {}) and d["cpu_baseline"].get("value"))
    except Exception as e:
        print(f, "parse failed", e)
PY
tail -9 gpurun_out/e2e_breakdown.txt; head -14 gpurun_out/launches.txt
