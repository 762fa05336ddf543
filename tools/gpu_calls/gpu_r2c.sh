#!/bin/bash
# round 2, GPU call C (2 GPUs): tests, slab bit-identity soak, N=1 and N=2 bench with PDL + overlapped upload on
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/r2c_tests.log 2>&1; echo "tests rc=$?" >> $O/r2c_tests.log
tail -4 $O/r2c_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/multigpu_check.py --size 4096 --csv-steps 12 --repeat 3 --trace > $O/r2c_mg2.log 2> $O/r2c_mg2.err; echo "mg2 rc=$?"; tail -2 $O/r2c_mg2.log
timeout 400 $TR tools/multigpu_check.py --size 16384 --square --tiles world --pm-T 5 --csv-steps 100 --repeat 3 --trace > $O/r2c_mg2_full.log 2> $O/r2c_mg2_full.err; echo "mg2 full rc=$?"; tail -2 $O/r2c_mg2_full.log
python bench.py --steps 5 --warmup 3 --no-extra --no-cpu > $O/r2c_bench_n1.json 2> $O/r2c_bench_n1.err; echo "bench n1 rc=$?"
CVB_BENCH_TRACE=1 timeout 400 $TR bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2c_bench_n2.json 2> $O/r2c_bench_n2.err; echo "bench n2 rc=$?"
echo done
