#!/bin/bash
# round 2, GPU call N (8 GPUs): final 8-GPU numbers -- soak of the bit-identity check, bench (slabs), bench (batch)
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
timeout 300 $TR --master-port 29801 tools/multigpu_check.py --size 4096 --csv-steps 12 --repeat 50 > $O/r2n_mg8_soak.log 2> $O/r2n_mg8_soak.err; echo "mg8 soak rc=$?"; tail -1 $O/r2n_mg8_soak.log
timeout 200 $TR --master-port 29802 bench.py --gpus 8 --steps 6 --warmup 3 > $O/r2n_bench_n8.json 2> $O/r2n_bench_n8.err; echo "bench n8 rc=$?"
timeout 300 $TR --master-port 29803 bench.py --gpus 8 --workload batch --steps 2 --warmup 1 > $O/r2n_batch_n8.json 2> $O/r2n_batch_n8.err; echo "batch n8 rc=$?"
echo done
