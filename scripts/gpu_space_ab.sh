#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-spaceab}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_space_memory.py tests/test_gpu_dropin_model.py -x -q -m gpu 2>&1 | tail -3 | cut -c1-300
for pf in 0 1 3; do echo "VADC_TC_PREFETCH=$pf"; VADC_TC_PREFETCH=$pf timeout -s KILL 300 python scripts/space_time.py 2>&1 | tail -1; done
echo "VADC_TC_TWO_STAGE=1 VADC_TC_PREFETCH=1"; VADC_TC_TWO_STAGE=1 VADC_TC_PREFETCH=1 timeout -s KILL 300 python scripts/space_time.py 2>&1 | tail -1
VADC_TC_PREFETCH=${BEST_PF:-1} timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/space_time.py 1 > gpurun_out/${TAG}_ncu.log 2>&1
python scripts/launch_metrics.py gpurun_out/${TAG}_launches.csv 30 | cut -c1-150
