"""encoder downsample stage (Conv3d 96->192 (1,2,2)/(1,2,2) + GELU -> channel-last tokens) at the cfg2 batch:
fused op vs torch's conv3d + gelu + the rearrange copy, forward + backward, median ms"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import statistics
import torch
import videoad_b200 as V
dev = torch.device("cuda", 0)
torch.manual_seed(0)
torch.backends.cudnn.allow_tf32 = False
seq = torch.nn.Sequential(torch.nn.Conv3d(96, 192, (1, 2, 2), stride=(1, 2, 2)), torch.nn.GELU()).to(dev)
B, D, H, W = 64, 8, 32, 32
x = torch.randn(B, 96, D, 2 * H, 2 * W, device=dev, requires_grad=True)
gy = torch.randn(B, D, H, W, 192, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(fn, n=6):
    ts = []
    for i in range(n):
        x.grad = None
        for p in seq.parameters(): p.grad = None
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y = fn(); y.backward(gy); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts[1:])


t_fused = run(lambda: V.downsample_gelu_tokens(x, seq[0], seq[1]))
t_torch = run(lambda: seq(x).permute(0, 2, 3, 4, 1).contiguous())
print(f"tokens {B * D * H * W}: fused fwd+bwd {t_fused:.3f} ms, torch conv3d+gelu+rearrange copy fwd+bwd {t_torch:.3f} ms ({t_torch / t_fused:.2f}x)")
