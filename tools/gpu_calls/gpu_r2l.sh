#!/bin/bash
# round 2, GPU call L (1 GPU): full tests, bench with extras, ncu launch list + full capture of csv_step (profiles/r2_*)
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -q ) > $O/r2l_tests.log 2>&1; echo "tests rc=$?" >> $O/r2l_tests.log; tail -3 $O/r2l_tests.log
python bench.py --steps 5 --warmup 3 > $O/r2l_bench.json 2> $O/r2l_bench.err; echo "bench rc=$?"
python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > $O/r2l_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/r2l_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > $O/r2l_ncu_launch.log 2>&1
python bench.py --size 8192 --steps 1 --warmup 1 --no-cpu --no-extra > $O/r2l_plain8k.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:csv_step_kernel -s 60 -c 1 -o $O/r2l_prof_csv -f \
    python bench.py --size 8192 --steps 1 --warmup 1 --no-cpu --no-extra > $O/r2l_ncu_full.log 2>&1
echo done
