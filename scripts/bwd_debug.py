"""which rows / outputs of the training-graph backward are wrong: python scripts/bwd_debug.py N [C]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import test_gpu_cluster as t
for n in [int(a) for a in sys.argv[1:]] or [1500]:
    got, ref = t._training_graph_backward(n, 192, 32, 16.0, seed=n + 192, scale_g=1e-2, loss_w=1.3)
    gx, rx = got[0], ref[0]
    err = np.abs(gx - rx).max(1) / np.abs(rx).max()
    bad = np.flatnonzero(err > 2e-4)
    print(f"N={n}: gx bad rows {bad.size}: {bad[:40]} ... max {err.max():.3e}; others:",
          [float(np.abs(g - r).max() / np.abs(r).max()) for g, r in zip(got[1:], ref[1:])], flush=True)
    if bad.size:
        r0 = bad[0]
        e = np.abs(gx[r0] - rx[r0]) / np.abs(rx).max()
        print("   first bad row", r0, "tile", r0 // 64, "row in tile", r0 % 64, "bad channels", np.flatnonzero(e > 2e-4)[:20], "n", (e > 2e-4).sum())
