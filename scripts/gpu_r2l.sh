#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 300 python scripts/bwd_stress.py 30 2>&1 | grep -v Warn | tail -3
bash scripts/gpu_ab.sh ${1:-r2l} 2>&1 | tail -3
D=video-anomaly-detection-guided-by-clustering-learning_b200
VADC_LIB_PATH=$PWD/$D/libvadc_trace.so VADC_BWD_TRACE=$PWD/gpurun_out/${1:-r2l}_trace.txt VADC_BWD_TRACE_CTA=17 timeout -s KILL 200 python scripts/bwd_only.py 3 > gpurun_out/${1:-r2l}_trace.log 2>&1
python scripts/trace_summary.py gpurun_out/${1:-r2l}_trace.txt 20 4 | tail -19
timeout -s KILL 600 python -m pytest tests/test_gpu_cluster.py -x -q -k "training_graph or golden" 2>&1 | tail -2
