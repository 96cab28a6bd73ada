"""one encoder-tail forward + backward at the cfg2 batch after a warm-up: for ncu launch lists"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import videoad_b200 as V
dev = torch.device("cuda", 0)
torch.manual_seed(0)
seq = torch.nn.Sequential(torch.nn.Conv3d(96, 192, (1, 2, 2), stride=(1, 2, 2)), torch.nn.GELU()).to(dev)
x = torch.randn(64, 96, 8, 64, 64, device=dev, requires_grad=True)
g = torch.randn(64, 8, 32, 32, 192, device=dev)
for i in range(2):
    x.grad = None
    for p in seq.parameters(): p.grad = None
    V.downsample_gelu_tokens(x, seq[0], seq[1]).backward(g)
torch.cuda.synchronize()
