#!/bin/bash
# one ncu --set full capture of the last launch of a kernel (regex $2) in a script ($3...), summary into gpurun_out/<tag>_ncu_prof.txt
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=$1; KERN=$2; shift 2
ncu --set full --clock-control none --import-source on -k regex:$KERN --launch-skip ${SKIP:-1} --launch-count 1 -o gpurun_out/prof_$TAG -f "$@" > gpurun_out/ncu_$TAG.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep 30 > gpurun_out/${TAG}_ncu_prof.txt 2>&1
python scripts/ncu_buckets.py gpurun_out/prof_$TAG.ncu-rep 60 >> gpurun_out/${TAG}_ncu_prof.txt 2>&1
head -30 gpurun_out/${TAG}_ncu_prof.txt
