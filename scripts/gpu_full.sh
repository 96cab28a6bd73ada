#!/bin/bash
# full GPU check: every -m gpu test, smoke, the bench line (N = 1)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-full}
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
tail -15 gpurun_out/${TAG}_pytest.log | cut -c1-300
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("h2d_GBps"))
print("dominant", r["kernel"], r["ms"], r["frac"], "other", r["other"]["ms"], r["other"]["frac"], "path", r["path"]["frac"], "step", r["step"]["frac"])
print("cpu", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("kind"), "gpu_ref", d.get("gpu_reference"))
for k, v in d.get("extra", {}).items(): print(" ", k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a != "timing"})
PY
