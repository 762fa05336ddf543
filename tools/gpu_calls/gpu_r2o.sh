#!/bin/bash
# round 2, GPU call O (1 GPU): final check -- all gpu tests, smoke, the default bench line with all extras
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -q ) > $O/r2o_tests.log 2>&1; echo "tests rc=$?" >> $O/r2o_tests.log; tail -3 $O/r2o_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2o_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2o_smoke.log
python bench.py > $O/r2o_bench.json 2> $O/r2o_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > $O/r2o_bench_ref.json 2> $O/r2o_bench_ref.err; echo "ref rc=$?"
echo done
