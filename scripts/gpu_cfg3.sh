#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout -s KILL 600 python -m pytest tests/test_gpu_cluster.py -x -q -m gpu 2>&1 | tail -2 | cut -c1-200
for K in 16 64 256; do
  timeout -s KILL 200 ncu --metrics $M --clock-control none -c 100 --csv --log-file gpurun_out/r2_cfg3_K$K.csv python scripts/fwd_generic.py 768 $K 65536 > gpurun_out/r2_cfg3_ncu.log 2>&1
  echo "C=768 K=$K"; python scripts/launch_metrics.py gpurun_out/r2_cfg3_K$K.csv 10 | cut -c1-150
done
