#!/bin/bash
# usage: scripts/prof_bwd.sh <tag> [kernel-regex]  -> gpurun_out/prof_<tag>.ncu-rep + per-launch time / DRAM bytes
TAG=${1:-x}; KRE=${2:-cluster_bwd_tc_kernel}
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:$KRE --launch-skip 3 --launch-count 2 --csv --log-file gpurun_out/launch_$TAG.csv python scripts/bwd_only.py 6 > gpurun_out/launch_$TAG.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launch_$TAG.csv")) if len(r)>5]
h=rows[0]; n=h.index("Metric Name"); v=h.index("Metric Value")
print("$TAG", [(r[n].split("__")[1][:12], round(float(r[v])/1e3,1)) for r in rows[1:]])
PY
ncu --set full --clock-control none --import-source on -k regex:$KRE --launch-skip 3 --launch-count 1 -o gpurun_out/prof_$TAG -f python scripts/bwd_only.py 6 > gpurun_out/ncu_$TAG.log 2>&1
