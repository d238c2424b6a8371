#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_densenet.py -k "one_piece" -q ) 2>&1 | tail -3
for dt in bf16 fp16; do
  ( timeout 600 python bench.py --dtype $dt --no-cpu-baseline ) > gpurun_out/bench_$dt.json 2> gpurun_out/bench_$dt.err; echo "bench $dt rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$dt.json"))
    print("$dt", {k: d[k] for k in ("value", "ms_per_step", "launches_per_step", "dtype")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel"], d["parity_check"]["rel_err"], d["final_loss"])
except Exception as e:
    print("bench parse failed", e)
PY
  tail -2 gpurun_out/bench_$dt.err
done
