#!/bin/bash
# round-2 profile evidence: bench line, ncu launch list of the same command, full captures of the two dominant kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 --no-extra --no-reference-legs > gpurun_out/r2_plain.json 2> gpurun_out/r2_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_step.csv python bench.py --steps 2 --warmup 3 --no-extra --no-graph --no-reference-legs > gpurun_out/r2_ncu_launch.log 2>&1
timeout -s KILL 600 bash scripts/prof_bwd.sh r2_bwdtc2 cluster_bwd_tc2_kernel
python scripts/ncu_summary.py gpurun_out/prof_r2_bwdtc2.ncu-rep 40 > gpurun_out/r2_ncu_prof_bwdtc2.txt 2>&1
python scripts/ncu_buckets.py gpurun_out/prof_r2_bwdtc2.ncu-rep 100 >> gpurun_out/r2_ncu_prof_bwdtc2.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:cluster_fwd_ws_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/prof_r2_fwdws -f python scripts/bwd_only.py 6 > gpurun_out/ncu_r2_fwdws.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_r2_fwdws.ncu-rep 25 > gpurun_out/r2_ncu_prof_fwdws.txt 2>&1
head -12 gpurun_out/r2_ncu_prof_bwdtc2.txt
python scripts/launch_shares.py gpurun_out/r2_launches_step.csv 2>/dev/null | head -30
