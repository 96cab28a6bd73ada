#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_cluster.py -x -q -k "training_graph or golden" > gpurun_out/r2f_pytest.log 2>&1
grep -E "^E |passed|failed|Error" gpurun_out/r2f_pytest.log | head -20
D=video-anomaly-detection-guided-by-clustering-learning_b200
VADC_LIB_PATH=$PWD/$D/libvadc_trace.so VADC_BWD_TRACE=$PWD/gpurun_out/r2f_trace.txt VADC_BWD_TRACE_CTA=17 timeout -s KILL 200 python scripts/bwd_only.py 3 > gpurun_out/r2f_trace.log 2>&1
python scripts/trace_summary.py gpurun_out/r2f_trace.txt 20 4 | tail -20
