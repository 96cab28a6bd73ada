"""cycles per tcgen05.mma (kind::f16 bf16, operands in shared memory) for the tile shapes the fused kernels consider"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videoad_b200 import _lib
out = torch.zeros(2, dtype=torch.int64, device="cuda")
l = _lib.lib()
reps = 64
print("  M    N  A   B    cyc/mma(total) cyc/mma(issue)  Mflop/mma  flop/cyc")
for (M, N, a, b) in [(128, 32, 0, 0), (128, 64, 0, 0), (128, 128, 0, 0), (128, 256, 0, 0), (128, 192, 0, 1), (128, 32, 1, 1), (128, 64, 1, 1),
                     (128, 192, 1, 1), (64, 32, 0, 0), (64, 64, 0, 0), (64, 192, 0, 1), (64, 32, 1, 1), (64, 192, 1, 1), (64, 256, 0, 0)]:
    for _ in range(2):
        _lib.check(l.vadc_debug_umma_bench(M, N, a, b, reps, _lib.ptr(out), _lib.stream()), "bench")
        torch.cuda.synchronize()
    tot, iss = [int(v) for v in out.cpu()]
    n = reps * 8
    fl = 2 * M * N * 16
    print(f"{M:4d} {N:4d}  {'MN' if a else 'K '}  {'MN' if b else 'K '}  {tot / n:10.1f} {iss / n:14.1f} {fl / 1e6:10.3f} {fl / (tot / n):9.0f}")
