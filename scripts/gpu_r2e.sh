#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; : > gpurun_out/ab_r2e.txt
for sl in 0 0; do
  VADC_BWD_SLEEP=$sl timeout -s KILL 200 python scripts/bwd_ab.py 20 2>&1 | grep -v Warn | sed "s/^/sleep=$sl /" >> gpurun_out/ab_r2e.txt
done
AB_IMPL=tc1 timeout -s KILL 200 python scripts/bwd_ab.py 20 2>&1 | grep -v Warn >> gpurun_out/ab_r2e.txt
cat gpurun_out/ab_r2e.txt
D=video-anomaly-detection-guided-by-clustering-learning_b200
VADC_LIB_PATH=$PWD/$D/libvadc_trace.so VADC_BWD_TRACE=$PWD/gpurun_out/r2e_trace.txt VADC_BWD_TRACE_CTA=17 timeout -s KILL 200 python scripts/bwd_only.py 3 > gpurun_out/r2e_trace.log 2>&1
python scripts/trace_summary.py gpurun_out/r2e_trace.txt 20 6 | tail -22
timeout -s KILL 600 python -m pytest tests/test_gpu_cluster.py -x -q -k "training_graph or golden" 2>&1 | tail -2
