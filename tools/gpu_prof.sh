#!/bin/bash
# profiling visit: one --set full capture of kernels whose base name matches $1 (regex), skipping $2 launches, $3 captured
mkdir -p gpurun_out
K=${1:-conv_dense}
S=${2:-24}
C=${3:-4}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C -f -o gpurun_out/prof_$K \
    python tools/profile_step.py --steps 2 > gpurun_out/prof_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/prof_$K.ncu-rep
