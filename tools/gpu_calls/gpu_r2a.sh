#!/bin/bash
# round 2, GPU call A: tests, bench, tile sweep, I256 variant A/B, C3 tile sweep, ncu launch list + full capture
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/r2a_smi.txt
( time python -m pytest tests -m gpu -x -q ) > $O/r2a_tests.log 2>&1; echo "tests rc=$?" >> $O/r2a_tests.log
tail -5 $O/r2a_tests.log
python bench.py --steps 5 --warmup 3 > $O/r2a_bench.json 2> $O/r2a_bench.err; echo "bench rc=$?"
for T in 80 186; do
  python bench.py --steps 3 --warmup 2 --no-extra --no-cpu --tile-rows $T > $O/r2a_bench_T$T.json 2> $O/r2a_bench_T$T.err; echo "bench T=$T rc=$?"
done
V=chan_vese_b200/lib/variants/libcvb_i256.so
CVB_LIB=$V python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not full_size" > $O/r2a_tests_i256.log 2>&1; echo "i256 tests rc=$?"
CVB_LIB=$V python bench.py --steps 3 --warmup 2 --no-extra --no-cpu > $O/r2a_bench_i256.json 2> $O/r2a_bench_i256.err; echo "bench i256 rc=$?"
for T in 40 48 59; do CVB_TILE_ROWS=$T python tools/bench_configs.py C3 >> $O/r2a_c3.txt 2>&1; done
python tools/bench_configs.py C1 C2 >> $O/r2a_c12.txt 2>&1
# ncu (after the same command lines have exited 0 above / here)
python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > $O/r2a_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2a_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > $O/r2a_ncu_launch.log 2>&1
python bench.py --size 8192 --steps 1 --warmup 1 --no-cpu --no-extra > $O/r2a_plain8k.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:csv_step_kernel -s 60 -c 1 -o $O/r2a_prof_csv -f \
    python bench.py --size 8192 --steps 1 --warmup 1 --no-cpu --no-extra > $O/r2a_ncu_full.log 2>&1
echo done
