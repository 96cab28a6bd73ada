#!/bin/bash
# ncu evidence for the tensor-bound distance GEMM (north_star): tc_gemm_kernel at C=768, K=256, N=524288 tokens
ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel --launch-skip 4 --launch-count 1 -o gpurun_out/prof_tcgemm -f python scripts/fwd_only.py auto 4 768 256 > gpurun_out/ncu_tcgemm.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launch_tcgemm.csv python scripts/fwd_only.py auto 4 768 256 > gpurun_out/launch_tcgemm.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launch_tcgemm.csv")) if len(r)>5]
h=rows[0]; k=h.index("Kernel Name"); v=h.index("Metric Value")
for r in rows[-12:]: print(r[k][:80], r[v])
PY
