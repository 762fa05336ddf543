#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
timeout 200 $TR --master-port 29701 bench.py --gpus 8 --steps 6 --warmup 3 > $O/r2j_bench_n8.json 2> $O/r2j_bench_n8.err; echo "bench n8 rc=$?"
nvidia-smi --query-gpu=index,clocks.sm,power.draw,power.limit --format=csv > $O/r2j_smi.txt
echo done
