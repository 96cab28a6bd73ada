#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-spaceprof}
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_space_memory.py -x -q -m gpu -k space 2>&1 | tail -3 | cut -c1-300
timeout -s KILL 300 python scripts/space_time.py 2>&1 | tail -1
SKIP=1 timeout -s KILL 300 bash scripts/prof_kernel.sh ${TAG}_lnB ln_bwd_transposed_tile_kernel python scripts/space_time.py 1 | head -60
SKIP=1 timeout -s KILL 300 bash scripts/prof_kernel.sh ${TAG}_gz "tc_gemm_kernel" python scripts/space_time.py 1 | head -40
