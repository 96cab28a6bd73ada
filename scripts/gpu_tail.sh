#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_encoder_tail.py -x -q -m gpu -s 2>&1 | tail -15 | cut -c1-300
timeout -s KILL 300 python scripts/tail_time.py 2>&1 | tail -4
