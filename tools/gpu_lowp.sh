#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_densenet.py -k "one_piece" -q ) > gpurun_out/pytest_lowp.txt 2>&1; echo "lowp rc=$?"
grep -E "^E  |passed|failed" gpurun_out/pytest_lowp.txt | head -40
( time timeout 1500 python -m pytest tests -m gpu -q -k "not one_piece" ) > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.txt
