#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout -s KILL 300 ncu --metrics $M --clock-control none -c 200 --csv --log-file gpurun_out/r2_mem_launches.csv python scripts/mem_one.py > gpurun_out/r2_mem_ncu.log 2>&1
python scripts/launch_metrics.py gpurun_out/r2_mem_launches.csv 17 | cut -c1-150
timeout -s KILL 300 ncu --metrics $M --clock-control none -c 300 --csv --log-file gpurun_out/r2_native_launches.csv python scripts/native_one.py > gpurun_out/r2_native_ncu.log 2>&1
python scripts/launch_metrics.py gpurun_out/r2_native_launches.csv 60 | cut -c1-150
