#!/bin/bash
# round 2, GPU call D (1 GPU): full gpu test suite, bench with extras, batch workload
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -q ) > $O/r2d_tests.log 2>&1; echo "tests rc=$?" >> $O/r2d_tests.log
tail -4 $O/r2d_tests.log
python bench.py --steps 5 --warmup 3 > $O/r2d_bench.json 2> $O/r2d_bench.err; echo "bench rc=$?"
python bench.py --workload batch --batch 512 --steps 2 --warmup 1 > $O/r2d_batch512.json 2> $O/r2d_batch512.err; echo "batch rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2d_smoke.log 2>&1; echo "smoke rc=$?"
echo done
