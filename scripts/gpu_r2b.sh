#!/bin/bash
# round 2, call B: where does the second-generation backward spend its time? event trace + ncu full capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
D=video-anomaly-detection-guided-by-clustering-learning_b200
VADC_LIB_PATH=$PWD/$D/libvadc_trace.so VADC_BWD_TRACE=$PWD/gpurun_out/r2b_trace.txt VADC_BWD_TRACE_CTA=17 timeout -s KILL 200 python scripts/bwd_only.py 3 > gpurun_out/r2b_trace.log 2>&1
tail -2 gpurun_out/r2b_trace.log
timeout -s KILL 200 python scripts/bwd_only.py 8 > gpurun_out/r2b_plain.log 2>&1; tail -1 gpurun_out/r2b_plain.log
timeout -s KILL 600 bash scripts/prof_bwd.sh r2b cluster_bwd_tc2_kernel
python scripts/ncu_summary.py gpurun_out/prof_r2b.ncu-rep 40 > gpurun_out/r2b_ncu_summary.txt 2>&1
python scripts/ncu_buckets.py gpurun_out/prof_r2b.ncu-rep 100 > gpurun_out/r2b_ncu_buckets.txt 2>&1
head -30 gpurun_out/r2b_ncu_summary.txt
