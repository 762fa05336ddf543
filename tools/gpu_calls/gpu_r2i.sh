#!/bin/bash
# round 2, GPU call I (2 GPUs): slab bit-identity with the new step tail, bench at 2 ranks
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/multigpu_check.py --size 4096 --csv-steps 12 --repeat 3 > $O/r2i_mg2.log 2> $O/r2i_mg2.err; echo "mg2 rc=$?"; tail -1 $O/r2i_mg2.log
CVB_SEG_MULT=2 timeout 300 $TR --master-port 29512 tools/multigpu_check.py --size 4096 --csv-steps 12 --tiles world > $O/r2i_mg2_m2.log 2> $O/r2i_mg2_m2.err; echo "mg2 m2 rc=$?"; tail -1 $O/r2i_mg2_m2.log
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2i_bench_n2.json 2> $O/r2i_bench_n2.err; echo "bench n2 rc=$?"
timeout 300 $TR --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2i_bench_n2b.json 2> $O/r2i_bench_n2b.err; echo "bench n2b rc=$?"
echo done
