#!/bin/bash
# round 2, GPU call G (8 GPUs): tile-length sweep of csv_step and pm2 at 8 ranks (how much of the per-launch loss is tail?)
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "segments_per_cta or config1 or config2 or early_stop" > $O/r2g_tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/r2g_tests.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
P=29600
for spec in "40 0 1" "64 0 1" "40 48 2"; do
  set -- $spec; P=$((P+1))
  CVB_SEG_MULT=$3 CVB_PM2_SEG_ROWS=$2 timeout 200 $TR --master-port $P bench.py --gpus 8 --steps 4 --warmup 2 --tile-rows $1 > $O/r2g_T$1_P$2_M$3.json 2> $O/r2g_T$1_P$2_M$3.err; echo "T=$1 PM2=$2 M=$3 rc=$?"
done
echo done
