#!/bin/bash
# round 2, GPU call P (4 GPUs): the final code at 2 and 4 ranks (scaling table, digest equality)
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 2 --master-port 29901 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2p_bench_n2.json 2> $O/r2p_bench_n2.err; echo "bench n2 rc=$?"
timeout 200 $TR --nproc-per-node 4 --master-port 29902 bench.py --gpus 4 --steps 5 --warmup 3 > $O/r2p_bench_n4.json 2> $O/r2p_bench_n4.err; echo "bench n4 rc=$?"
echo done
