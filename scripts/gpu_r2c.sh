#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
bash scripts/gpu_ab.sh r2c
D=video-anomaly-detection-guided-by-clustering-learning_b200
VADC_LIB_PATH=$PWD/$D/libvadc_trace.so VADC_BWD_TRACE=$PWD/gpurun_out/r2c_trace.txt VADC_BWD_TRACE_CTA=17 timeout -s KILL 200 python scripts/bwd_only.py 3 > gpurun_out/r2c_trace.log 2>&1
python scripts/trace_summary.py gpurun_out/r2c_trace.txt 20 6 | tail -24
