#!/bin/bash
# round 2, GPU call B: PM temporal blocking (tests, bench, ncu), register-budget variants of csv_step
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/r2b_tests.log 2>&1; echo "tests rc=$?" >> $O/r2b_tests.log
tail -6 $O/r2b_tests.log
python bench.py --steps 5 --warmup 3 --no-extra --no-cpu > $O/r2b_bench.json 2> $O/r2b_bench.err; echo "bench rc=$?"
CVB_PM_FUSE=0 python bench.py --steps 3 --warmup 2 --no-extra --no-cpu > $O/r2b_bench_nofuse.json 2> $O/r2b_bench_nofuse.err; echo "bench nofuse rc=$?"
for v in c12 c20 pm2c16; do
  CVB_LIB=chan_vese_b200/lib/variants/libcvb_$v.so python bench.py --steps 3 --warmup 2 --no-extra --no-cpu > $O/r2b_bench_$v.json 2> $O/r2b_bench_$v.err; echo "bench $v rc=$?"
done
python tools/bench_configs.py C1 C2 C3 > $O/r2b_c123.txt 2>&1
python bench.py --size 8192 --steps 1 --warmup 1 --no-cpu --no-extra > $O/r2b_plain8k.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pm2_step_kernel -s 10 -c 1 -o $O/r2b_prof_pm2 -f \
    python bench.py --size 8192 --steps 1 --warmup 1 --no-cpu --no-extra > $O/r2b_ncu_full.log 2>&1
echo done
