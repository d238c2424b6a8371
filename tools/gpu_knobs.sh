#!/bin/bash
# knob sweep of the step time (bench.py value / e2e, no CPU baselines)
mkdir -p gpurun_out
run() {
  echo "== $*"
  env "$@" timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['stock_adam']['ms_per_step'])"
}
run PDES_DUMMY=1
run PDES_PDL=0
run PDES_PDL=0 PDES_WGRAD_STREAMS=0
run PDES_BENCH_GRAPH=0
run PDES_BENCH_GRAPH=0 PDES_WGRAD_STREAMS=0
