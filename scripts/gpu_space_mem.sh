#!/bin/bash
# space head + memory: parity tests, timings, launch list of the space step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-spacemem}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_space_memory.py tests/test_gpu_dropin_model.py tests/test_gpu_cluster.py -x -q -m gpu 2>&1 | tail -12 | cut -c1-300
for t in 2 3; do echo "VADC_SPACE_TERMS=$t"; VADC_SPACE_TERMS=$t timeout -s KILL 300 python scripts/space_time.py 2>&1 | tail -1; done
for t in 2 3; do echo "VADC_MEMORY_TERMS=$t"; VADC_MEMORY_TERMS=$t timeout -s KILL 300 python scripts/memory_time.py 2>&1 | tail -3; done
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/space_time.py 1 > gpurun_out/${TAG}_ncu.log 2>&1
python scripts/launch_metrics.py gpurun_out/${TAG}_launches.csv 28 | cut -c1-150
