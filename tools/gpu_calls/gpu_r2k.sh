#!/bin/bash
# round 2, GPU call K (1 GPU): tests, tile 123 vs 128 on one box, bench with extras
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -q ) > $O/r2k_tests.log 2>&1; echo "tests rc=$?" >> $O/r2k_tests.log; tail -3 $O/r2k_tests.log
for T in 123 128 124 126; do
python bench.py --steps 4 --warmup 2 --no-extra --no-cpu --tile-rows $T > $O/r2k_bench_T$T.json 2> $O/r2k_bench_T$T.err; echo "bench T=$T rc=$?"
done
echo done
