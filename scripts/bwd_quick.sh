#!/bin/bash
# parity (tc vs mma.sync vs oracle) + kernel time of the tcgen05 backward
timeout 120 python scripts/bwd_check.py 2>&1 | tail -6 | cut -c1-110
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"cluster_bwd_tc_kernel" --launch-skip 3 --launch-count 3 python scripts/bwd_only.py 6 2>&1 | grep -E "duration"
