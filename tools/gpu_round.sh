#!/bin/bash
# One GPU-box visit: parity tests, the bench line, in-situ per-launch timing, tc2 phase timeline.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"
( time timeout 900 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python tools/layer_timing.py > gpurun_out/layer_timing.out 2> gpurun_out/layer_timing.txt; echo "timing rc=$?"
timeout 300 python tools/tc2_timeline.py > gpurun_out/tc2_timeline.out 2> gpurun_out/tc2_timeline.txt; echo "timeline rc=$?"
tail -3 gpurun_out/pytest_gpu.txt; cat gpurun_out/bench.json
