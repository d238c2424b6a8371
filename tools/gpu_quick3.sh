#!/bin/bash
# short visit: conv / densenet parity, bench, fused-dgrad phase stamps
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_conv.py tests/test_gpu_densenet.py -m gpu -q -x ) > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.txt
( time timeout 900 python bench.py --no-cpu-baseline ) > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "launches_per_step")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline_families"], d["roofline_stencil"]["frac"], d["parity_check"]["rel_err"])
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench.err
PDES_DENSE_DBG_BWD=${1:-6} timeout 300 python tools/profile_step.py --steps 3 2>&1 | grep -i "CTA\|loss" | head -8
PDES_WGRAD_STREAMS=0 timeout 300 python tools/layer_timing.py > gpurun_out/layer_timing.out 2> gpurun_out/layer_timing_s0.txt; echo "timing rc=$?"
python tools/timing_summary.py gpurun_out/layer_timing_s0.txt 2>/dev/null | head -16
