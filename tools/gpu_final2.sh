#!/bin/bash
# round-end evidence: tests, smoke, bench (own arm + CPU arm + bf16), script/solver on the GPU, launch list, ncu captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; grep -n "passed\|failed" gpurun_out/pytest_gpu.txt | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.txt
( time timeout 900 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
( time timeout 600 python bench.py --impl reference --steps 10 --warmup 2 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
( timeout 600 python bench.py --dtype bf16 --no-cpu-baseline ) > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bf16 rc=$?"
( timeout 600 python bench.py --dtype fp16 --no-cpu-baseline ) > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; echo "fp16 rc=$?"
( timeout 600 python tools/run_script_gpu.py --epochs 4 --out gpurun_out/script_run ) > gpurun_out/script_run.out 2>&1; echo "script rc=$?"
( timeout 600 python tools/run_solver_gpu.py --epochs 25 ) > gpurun_out/solver_run.out 2>&1; echo "solver rc=$?"
timeout 300 python tools/e2e_breakdown.py > gpurun_out/e2e_breakdown.txt 2>&1; echo "e2e breakdown rc=$?"
export PDES_EXEC_GRAPH=0
PDES_WGRAD_STREAMS=0 timeout 300 python tools/layer_timing.py > gpurun_out/layer_timing.out 2> gpurun_out/layer_timing_s0.txt; echo "timing rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py --steps 3 > gpurun_out/prof_step.log 2>&1; echo "list rc=$?"
python tools/launch_summary.py gpurun_out/launches.csv 3 30 > gpurun_out/launches.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:'conv_dense_fwd_kernel|act_split_staged_kernel' -s 27 -c 8 -o gpurun_out/prof_fwd python tools/profile_step.py --steps 2 > gpurun_out/prof_fwd.log 2>&1; echo "ncu fwd rc=$?"
timeout 600 $NCU -k regex:'conv_dense_bwd_kernel|wgrad_tn_kernel' -s 40 -c 8 -o gpurun_out/prof_bwd python tools/profile_step.py --steps 2 > gpurun_out/prof_bwd.log 2>&1; echo "ncu bwd rc=$?"
timeout 600 $NCU -k regex:conv_tc2_kernel -s 24 -c 6 -o gpurun_out/prof_tc2 python tools/profile_step.py --steps 2 > gpurun_out/prof_tc2.log 2>&1; echo "ncu tc2 rc=$?"
unset PDES_EXEC_GRAPH
ls -la gpurun_out/*.ncu-rep
python - <<'PY'
import json
for f in ("bench", "bench_bf16", "bench_fp16"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, {k: d[k] for k in ("value", "ms_per_step", "launches_per_step", "dtype")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"],
              "roof", d["roofline"]["frac"], d["roofline_step"]["frac"], "stencil", d["roofline_stencil"]["frac"], d.get("gpu_library_baseline"), d.get("cpu_baseline", {}) and d["cpu_baseline"].get("value"))
    except Exception as e:
        print(f, "parse failed", e)
PY
cat gpurun_out/bench_ref.json | head -c 600; echo; tail -c 700 gpurun_out/script_run.out; echo; tail -c 400 gpurun_out/solver_run.out; echo; cat gpurun_out/e2e_breakdown.txt | tail -9; head -12 gpurun_out/launches.txt
