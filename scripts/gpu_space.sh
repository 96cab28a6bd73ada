#!/bin/bash
# space head: parity tests, then the launch list of one forward + backward at the cfg2 batch
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-space}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_space_memory.py tests/test_gpu_dropin_model.py -x -q -m gpu 2>&1 | tail -15 | cut -c1-300
timeout -s KILL 300 python scripts/space_time.py 2>&1 | tail -8
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/space_time.py 1 > gpurun_out/${TAG}_ncu.log 2>&1
python scripts/launch_metrics.py gpurun_out/${TAG}_launches.csv 30
