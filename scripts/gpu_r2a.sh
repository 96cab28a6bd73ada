#!/bin/bash
# round 2, call A: first run of the second-generation backward (bt2) + the new parity tests + same-box A/B bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2a_gpu.txt 2>&1
# 1. quick smoke of the new kernel under a hard timeout (a deadlock must not eat the call)
timeout -s KILL 180 python - > gpurun_out/r2a_quick.log 2>&1 <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import test_gpu_cluster as t
for (n, C) in ((64, 192), (1, 192), (1500, 192), (777, 128), (333, 64), (148 * 64 * 3 + 5, 192)):
    got, ref = t._training_graph_backward(n, C, 32, 16.0, seed=n + C, scale_g=1e-2, loss_w=1.3)
    torch.cuda.synchronize()
    print(n, C, [float(np.abs(g - r).max() / np.abs(r).max()) for g, r in zip(got, ref)], flush=True)
PY
echo "quick rc=$?" >> gpurun_out/r2a_quick.log
tail -8 gpurun_out/r2a_quick.log
# 2. the cluster parity suite (includes the new full-size backward check)
timeout -s KILL 900 python -m pytest tests/test_gpu_cluster.py -x -q > gpurun_out/r2a_pytest_cluster.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest_cluster.log
tail -5 gpurun_out/r2a_pytest_cluster.log
# 3. same-box A/B: first generation vs second generation
for impl in tc1 tc; do
  VADC_BWD_IMPL=$impl timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --no-extra > gpurun_out/r2a_bench_$impl.json 2> gpurun_out/r2a_bench_$impl.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2a_bench_$impl.json").read().strip().splitlines()[-1])
    r = d["roofline"]; o = r.get("other", {})
    print("$impl", "ms/step", round(d["ms_per_step"], 4), "| dominant", r["kernel"][:22], round(r["ms"], 4), round(r["frac"], 3), "| other", round(o.get("ms", 0), 4), round(o.get("frac", 0), 3), "| path", round(r["path"]["frac"], 3))
except Exception as e:
    print("$impl", "bench failed", e)
PY
done
