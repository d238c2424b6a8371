for w in 1 2; do echo "== waves $w"; PDES_WG_WAVES=$w timeout 300 python bench.py --no-cpu-baseline --steps 50 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'conv', d['conv_path_ms_per_step'], 'loss', d['final_loss'])"
PDES_WG_WAVES=$w PDES_WGRAD_STREAMS=0 timeout 300 python tools/layer_timing.py > gpurun_out/layer_timing.out 2> gpurun_out/layer_timing_s0.txt
python tools/timing_summary.py gpurun_out/layer_timing_s0.txt 2>/dev/null | head -4; done
