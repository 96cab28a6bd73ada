#!/bin/bash
# multi-GPU bench line: scripts/gpu_multi.sh <N> <tag>
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}; TAG=${2:-multi}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
tail -5 gpurun_out/${TAG}_bench_${N}gpu.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_${N}gpu.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "ms/step", d["ms_per_step"], "value", d["value"], "launch", d["launch_path"][:40])
print("dp_parity", d.get("dp_parity"))
print("strong", d.get("strong")); print("weak", d.get("weak"))
print("e2e", d["e2e"]["value"], d["e2e"]["h2d_GBps"])
PY
