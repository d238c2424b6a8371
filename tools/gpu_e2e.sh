#!/bin/bash
mkdir -p gpurun_out
echo "== fused adam"; timeout 300 python tools/e2e_breakdown.py 2>&1 | tail -9
echo "== exec graph"; PDES_EXEC_GRAPH=1 timeout 300 python tools/e2e_breakdown.py 2>&1 | tail -9
