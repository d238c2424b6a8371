#!/bin/bash
# profiling visit: launch list of a step + one --set full capture of kernels matching $1 (regex), skipping $2 launches
mkdir -p gpurun_out
K=${1:-conv_dense}
S=${2:-24}
C=${3:-4}
#timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py --steps 3 > gpurun_out/prof_step.log 2>&1; echo "list rc=$?"
#python tools/launch_summary.py gpurun_out/launches.csv 3 40 > gpurun_out/launches.txt 2>&1; head -45 gpurun_out/launches.txt
timeout 900 ncu --set full --cache-control none --clock-control none --import-source on -k regex:$K -s $S -c $C -f -o gpurun_out/prof_$K \
    python tools/profile_step.py --steps 2 > gpurun_out/prof_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/*.ncu-rep
