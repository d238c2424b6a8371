#!/bin/bash
# One GPU-box visit (round 2): parity tests (not -x: see every failure), smoke, bench line, the unmodified script.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.txt
( time timeout 900 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
( time timeout 600 python tools/run_script_gpu.py --epochs 3 --out gpurun_out/script_run ) > gpurun_out/script_run.out 2>&1; echo "script rc=$?"
tail -15 gpurun_out/pytest_gpu.txt; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; tail -2 gpurun_out/script_run.out
