#!/bin/bash
# round 2, GPU call F (8 GPUs): slab bit-identity at 8 ranks, bench at 8 / 4 ranks (PDL + overlapped upload on), batch workload
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29511 tools/multigpu_check.py --size 4096 --csv-steps 12 --repeat 3 --trace > $O/r2f_mg8.log 2> $O/r2f_mg8.err; echo "mg8 rc=$?"; tail -1 $O/r2f_mg8.log
for rep in 1 2 3; do
CVB_BENCH_TRACE=1 timeout 300 $TR --nproc-per-node 8 --master-port 2952$rep bench.py --gpus 8 --steps 5 --warmup 3 > $O/r2f_bench_n8_$rep.json 2> $O/r2f_bench_n8_$rep.err; echo "bench n8 #$rep rc=$?"
done
timeout 300 $TR --nproc-per-node 4 --master-port 29531 bench.py --gpus 4 --steps 5 --warmup 3 > $O/r2f_bench_n4.json 2> $O/r2f_bench_n4.err; echo "bench n4 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --workload batch --steps 2 --warmup 1 > $O/r2f_batch_n8.json 2> $O/r2f_batch_n8.err; echo "batch n8 rc=$?"
python bench.py --steps 5 --warmup 3 --no-extra --no-cpu > $O/r2f_bench_n1.json 2> $O/r2f_bench_n1.err; echo "bench n1 rc=$?"
echo done
