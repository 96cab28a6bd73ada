#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; : > gpurun_out/ab_r2p.txt
for pf in 0 1 2; do
  VADC_BWD_PF=$pf timeout -s KILL 200 python scripts/bwd_ab.py 20 2>&1 | grep -v Warn | sed "s/^/pf=$pf /" >> gpurun_out/ab_r2p.txt
done
cat gpurun_out/ab_r2p.txt
D=video-anomaly-detection-guided-by-clustering-learning_b200
VADC_BWD_PF=0 VADC_LIB_PATH=$PWD/$D/libvadc_trace.so VADC_BWD_TRACE=$PWD/gpurun_out/r2p_trace.txt VADC_BWD_TRACE_CTA=17 timeout -s KILL 200 python scripts/bwd_only.py 3 > gpurun_out/r2p_trace.log 2>&1
python scripts/trace_summary.py gpurun_out/r2p_trace.txt 20 4 | tail -19
