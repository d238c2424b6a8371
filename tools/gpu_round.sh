#!/bin/bash
# One GPU-box visit: parity tests, smoke, the bench line, in-situ per-launch timing.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.txt
( time timeout 900 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
( time timeout 600 python bench.py --impl reference --steps 10 --warmup 2 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
PDES_WGRAD_STREAMS=0 timeout 300 python tools/layer_timing.py > gpurun_out/layer_timing.out 2> gpurun_out/layer_timing_s0.txt; echo "timing rc=$?"
tail -3 gpurun_out/pytest_gpu.txt; cat gpurun_out/bench.json; cat gpurun_out/bench_ref.json
