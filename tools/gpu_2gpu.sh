#!/bin/bash
# two ranks: bench (value + e2e with the fused Adam and executor graphs under DDP) and the gradient-bucket parity check
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"
tail -c 900 gpurun_out/bench_2gpu.json; echo; tail -3 gpurun_out/bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    tools/ddp_parity.py > gpurun_out/ddp_parity.json 2> gpurun_out/ddp_parity.err; echo "parity rc=$?"
tail -c 600 gpurun_out/ddp_parity.json; echo; tail -3 gpurun_out/ddp_parity.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_2gpu.json 2> gpurun_out/bench_ref_2gpu.err; echo "ref2 rc=$?"
tail -c 300 gpurun_out/bench_ref_2gpu.json
