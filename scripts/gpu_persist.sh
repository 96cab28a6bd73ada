#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests/test_gpu_tc_gemm.py -x -q -m gpu 2>&1 | tail -4 | cut -c1-300
timeout -s KILL 900 python -m pytest tests/test_gpu_space_memory.py tests/test_gpu_cluster.py tests/test_gpu_decoder_entry.py tests/test_gpu_encoder_tail.py tests/test_gpu_dropin_model.py -x -q -m gpu 2>&1 | tail -4 | cut -c1-300
timeout -s KILL 200 python scripts/memory_time.py | tail -3
VADC_TC_NO_PERSIST=1 timeout -s KILL 200 python scripts/memory_time.py | tail -1
timeout -s KILL 200 python scripts/tail_time.py | tail -1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout -s KILL 300 ncu --metrics $M --clock-control none -c 300 --csv --log-file gpurun_out/r2_native_launches.csv python scripts/native_one.py fused > gpurun_out/r2_native_ncu.log 2>&1
python scripts/launch_metrics.py gpurun_out/r2_native_launches.csv 42 | grep "tc_gemm\|total" | cut -c1-150
timeout -s KILL 300 ncu --metrics $M --clock-control none -c 200 --csv --log-file gpurun_out/r2_mem_launches.csv python scripts/mem_one.py > gpurun_out/r2_mem_ncu.log 2>&1
python scripts/launch_metrics.py gpurun_out/r2_mem_launches.csv 17 | grep "tc_gemm\|total" | cut -c1-150
