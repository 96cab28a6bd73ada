#!/bin/bash
# same-box A/B of backward-kernel variants: scripts/gpu_ab.sh <tag> [lib.so ...]   (default library first)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=$1; shift
mkdir -p gpurun_out
: > gpurun_out/ab_$TAG.txt
timeout -s KILL 200 python scripts/bwd_ab.py 20 >> gpurun_out/ab_$TAG.txt 2>&1
for L in "$@"; do
  VADC_LIB_PATH=$PWD/$L timeout -s KILL 200 python scripts/bwd_ab.py 20 >> gpurun_out/ab_$TAG.txt 2>&1
done
timeout -s KILL 200 python scripts/bwd_ab.py 20 >> gpurun_out/ab_$TAG.txt 2>&1
grep -v Warning gpurun_out/ab_$TAG.txt
