#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_glow.py -q ) > gpurun_out/pytest_glow.txt 2>&1; echo "glow rc=$?"
grep -E "^E  |passed|failed|FAILED" gpurun_out/pytest_glow.txt | head -30
( time timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_glow.py ) > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.txt
