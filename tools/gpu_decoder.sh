#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_densenet.py -k "decoder" -q ) > gpurun_out/pytest_decoder.txt 2>&1; echo "decoder rc=$?"
grep -E "^E  |passed|failed" gpurun_out/pytest_decoder.txt | head -30
( timeout 600 python tools/run_solver_gpu.py --epochs 25 ) > gpurun_out/solver_run.out 2>&1; echo "solver rc=$?"; tail -c 1500 gpurun_out/solver_run.out
( time timeout 1500 python -m pytest tests -m gpu -q -k "not decoder" ) > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.txt
