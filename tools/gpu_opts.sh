#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_densenet.py -k "dropout or bilinear" -q ) > gpurun_out/pytest_opts.txt 2>&1; echo "opts rc=$?"
grep -E "^E  |passed|failed|FAILED" gpurun_out/pytest_opts.txt | head -30
( time timeout 1500 python -m pytest tests -m gpu -q -k "not dropout and not bilinear" ) > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.txt
