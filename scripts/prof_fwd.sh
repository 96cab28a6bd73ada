#!/bin/bash
# usage: scripts/prof_fwd.sh <tag> [kernel-regex]  -> gpurun_out/prof_<tag>.ncu-rep + launch list
TAG=${1:-x}; KRE=${2:-cluster_fwd_ws_kernel}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launch_$TAG.csv python scripts/fwd_only.py auto 6 > gpurun_out/launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$KRE --launch-skip 3 --launch-count 1 -o gpurun_out/prof_$TAG -f python scripts/fwd_only.py auto 6 > gpurun_out/ncu_$TAG.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launch_$TAG.csv")) if len(r)>5]
h=rows[0]; k=h.index("Kernel Name"); v=h.index("Metric Value")
for r in rows[-9:]: print(r[k][:70], r[v])
PY
