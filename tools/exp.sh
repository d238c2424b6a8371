timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline --steps 50 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'conv', d['conv_path_ms_per_step'], 'loss', d['final_loss'])"
PDES_WGRAD_STREAMS=0 timeout 300 python tools/layer_timing.py > gpurun_out/layer_timing.out 2> gpurun_out/layer_timing_s0.txt
python tools/timing_summary.py gpurun_out/layer_timing_s0.txt 2>/dev/null | head -${1:-16}
