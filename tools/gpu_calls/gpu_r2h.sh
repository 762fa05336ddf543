#!/bin/bash
# round 2, GPU call H (1 GPU): full tests; the TMA build variant: parity tests, bench, ncu of both row loops
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -q ) > $O/r2h_tests.log 2>&1; echo "tests rc=$?" >> $O/r2h_tests.log; tail -3 $O/r2h_tests.log
V=chan_vese_b200/lib/variants/libcvb_tma.so
CVB_LIB=$V timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $O/r2h_tests_tma.log 2>&1; echo "tma tests rc=$?"; tail -3 $O/r2h_tests_tma.log
python bench.py --steps 5 --warmup 3 --no-extra --no-cpu > $O/r2h_bench.json 2> $O/r2h_bench.err; echo "bench rc=$?"
CVB_LIB=$V timeout 300 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu > $O/r2h_bench_tma.json 2> $O/r2h_bench_tma.err; echo "bench tma rc=$?"
python bench.py --steps 5 --warmup 3 --no-extra --no-cpu > $O/r2h_bench2.json 2> $O/r2h_bench2.err; echo "bench2 rc=$?"
CVB_LIB=$V python bench.py --size 8192 --steps 1 --warmup 1 --no-cpu --no-extra > $O/r2h_plain8k.log 2>&1 &&
CVB_LIB=$V ncu --set full --clock-control none --import-source on -k regex:csv_step_kernel -s 60 -c 1 -o $O/r2h_prof_csv_tma -f \
    python bench.py --size 8192 --steps 1 --warmup 1 --no-cpu --no-extra > $O/r2h_ncu_tma.log 2>&1
echo done
