#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_space_memory.py tests/test_gpu_cluster.py tests/test_gpu_dropin_model.py -x -q -m gpu 2>&1 | tail -4 | cut -c1-300
for e in 0 1; do echo "VADC_TC_NO_PERSIST_BATCHED=$e"; VADC_TC_NO_PERSIST_BATCHED=$e timeout -s KILL 300 python scripts/space_time.py 2>&1 | tail -1; done
VADC_SPACE_TERMS=3 timeout -s KILL 300 python scripts/space_time.py 2>&1 | tail -1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2space11_launches.csv python scripts/space_time.py 1 > gpurun_out/r2space11_ncu.log 2>&1
python scripts/launch_metrics.py gpurun_out/r2space11_launches.csv 28 | grep "tc_gemm\|total" | cut -c1-150
