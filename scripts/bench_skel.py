"""Isolated soft-skeleton timing at the train step's shape (N x 128^3, iters 15): python scripts/bench_skel.py [N]"""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from van_gan_b200 import clDice_func as K
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S, iters = 128, 15
x = torch.rand((N, S, S, S, 1), device="cuda")
g = torch.randn_like(x)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
def timed(fn):
    for _ in range(2): fn()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[2]
tf = timed(lambda: K.soft_skel(x, iters))
skel, bwd = K.soft_skel_with_grad(x, iters)
tb = timed(lambda: bwd(g))
model = 16.0 * x.numel() * (iters + 1)
print("soft_skel N=%d S=%d iters=%d: fwd %.3f ms (%.0f GB/s of the 16 B*V*(k+1) model), bwd %.3f ms (%.1fx fwd)" % (N, S, iters, tf, model / tf / 1e6, tb, tb / tf))
