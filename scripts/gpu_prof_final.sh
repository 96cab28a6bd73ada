#!/bin/bash
# full ncu captures of the round-2 secondary kernels (summaries for profiles/)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
SKIP=1 timeout -s KILL 300 bash scripts/prof_kernel.sh r2_persist_logits "tc_gemm_persist_kernel" python scripts/mem_one.py | head -32
SKIP=1 timeout -s KILL 300 bash scripts/prof_kernel.sh r2_lnT ln_transpose_tile_kernel python scripts/space_time.py 1 | head -16
SKIP=1 timeout -s KILL 300 bash scripts/prof_kernel.sh r2_lnB ln_bwd_transposed_tile_kernel python scripts/space_time.py 1 | head -16
