#!/bin/bash
# sweep the forward kernel's L2 prefetch distance / store cache hints: kernel time + DRAM bytes per launch (ncu)
for CFG in "1 0" "1 3" "-1 0" "-1 3" "-2 3" "-3 3" "2 3" "0 3"; do
  set -- $CFG
  VADC_WS_PF=$1 VADC_WS_HINT=$2 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:cluster_fwd_ws_kernel --launch-skip 3 --launch-count 2 --csv --log-file gpurun_out/pf.csv python scripts/fwd_only.py auto 6 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/pf.csv")) if len(r)>5]
h=rows[0]; n=h.index("Metric Name"); v=h.index("Metric Value")
print("PF=$1 HINT=$2", [(r[n].split("__")[1][:12], round(float(r[v])/1e3,1)) for r in rows[1:]])
PY
done
