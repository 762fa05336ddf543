#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -q ) > $O/r2e_tests.log 2>&1; echo "tests rc=$?" >> $O/r2e_tests.log
tail -4 $O/r2e_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2e_smoke.log 2>&1; echo "smoke rc=$?"
echo done
