#!/bin/bash
# parity + bench + compute-sanitizer memcheck over the small cases (new kernels: act_split staged, wgrad_tn, G ring, convT)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_conv.py tests/test_gpu_densenet.py -m gpu -q -x ) > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"
grep -n "passed\|failed" gpurun_out/pytest_gpu.txt | tail -2
( time timeout 900 python bench.py --no-cpu-baseline ) > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "launches_per_step")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline_families"], d["roofline_stencil"]["frac"], d["parity_check"]["rel_err"])
except Exception as e:
    print("bench parse failed", e)
PY
PDES_DENSE_DBG_BWD=6 timeout 300 python tools/profile_step.py --steps 2 2>&1 | grep -i "CTA" | head -2
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_densenet.py tests/test_gpu_conv.py -m gpu -q -x \
    -k "small16 or convt16 or bilinear16 or fiveblk16 or dropout or dense" ) > gpurun_out/sanitizer_memcheck.txt 2>&1; echo "sanitizer rc=$?"
tail -4 gpurun_out/sanitizer_memcheck.txt
