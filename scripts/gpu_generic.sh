#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_space_memory.py tests/test_gpu_dropin_model.py tests/test_gpu_decoder_entry.py -x -q -m gpu 2>&1 | tail -5 | cut -c1-300
python scripts/memory_time.py | tail -3
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout -s KILL 300 ncu --metrics $M --clock-control none -c 300 --csv --log-file gpurun_out/r2_native_launches.csv python scripts/native_one.py fused > gpurun_out/r2_native_ncu.log 2>&1
python scripts/launch_metrics.py gpurun_out/r2_native_launches.csv 42 | cut -c1-150
