#!/bin/bash
# round 2, GPU call M (1 GPU): full tests (fp32 ring, C3 2000 steps, C2 full field), bench with extras,
# ncu --set full of csv_step at the bench's own size (16384^2) for roofline.traffic
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -q ) > $O/r2m_tests.log 2>&1; echo "tests rc=$?" >> $O/r2m_tests.log; tail -3 $O/r2m_tests.log
python bench.py --steps 5 --warmup 3 > $O/r2m_bench.json 2> $O/r2m_bench.err; echo "bench rc=$?"
python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > $O/r2m_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:csv_step_kernel -s 120 -c 1 -o $O/r2m_prof_csv16k -f \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > $O/r2m_ncu_full.log 2>&1
echo done
