#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-spaceprof}
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_space_memory.py -x -q -m gpu -k space 2>&1 | tail -3 | cut -c1-300
timeout -s KILL 300 python scripts/space_time.py 2>&1 | tail -3
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/space_time.py 1 > gpurun_out/${TAG}_ncu.log 2>&1
python scripts/launch_metrics.py gpurun_out/${TAG}_launches.csv 30 | cut -c1-150
SKIP=1 timeout -s KILL 300 bash scripts/prof_kernel.sh ${TAG}_lnT ln_transpose_tile_kernel python scripts/space_time.py 1 | head -36
SKIP=1 timeout -s KILL 300 bash scripts/prof_kernel.sh ${TAG}_lnB ln_bwd_transposed_tile_kernel python scripts/space_time.py 1 | head -36
SKIP=1 timeout -s KILL 300 bash scripts/prof_kernel.sh ${TAG}_gz TcSpaceGzEpi python scripts/space_time.py 1 | head -36
