#!/bin/bash
# Tuning sweep on the GPU box: tools/sweep.sh "<nvcc flags>|<bench args>" ...   (results: gpurun_out/sweep.log)
mkdir -p gpurun_out
for spec in "$@"; do
  flags="${spec%%|*}"; bargs="${spec#*|}"
  CVB_EXTRA_NVCC_FLAGS="$flags" python -m chan_vese_b200.build --force > /dev/null 2>&1 || { echo "BUILD FAILED: $flags" | tee -a gpurun_out/sweep.log; continue; }
  out=$(timeout 300 python bench.py --no-cpu --steps 2 --warmup 1 $bargs 2>&1 | tail -1)
  echo "$spec => $(python - "$out" <<'PY'
import json,sys
try:
    d=json.loads(sys.argv[1]); k=d["kernels"]
    print("value=%.3e ms/step=%.1f csv=%.3f ms (%.3f) pm=%.3f ms (%.3f) sm=%s %s" % (d["value"], d["ms_per_step"], k["csv_step"]["ms_per_launch"], k["csv_step"]["frac"], k["pm_step"]["ms_per_launch"], k["pm_step"]["frac"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e:
    print("ERR", sys.argv[1][-300:])
PY
)" | tee -a gpurun_out/sweep.log
done
python -m chan_vese_b200.build --force > /dev/null 2>&1
