"""Isolated conv timing: python scripts/bench_conv.py [fwd|dgrad|wgrad] — prints TFLOP/s per layer shape."""
import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collections import OrderedDict
from van_gan_b200 import engine as E, _lib
from van_gan_b200._lib import call
mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
only = sys.argv[2] if len(sys.argv) > 2 else None
# (Cin, Cout, K, stride, S_in(padded), N)
shapes = [(16, 16, 3, 1, 130, 8), (48, 16, 3, 1, 130, 8), (32, 32, 3, 1, 66, 8), (96, 32, 3, 1, 66, 8), (64, 64, 3, 1, 34, 8),
          (192, 64, 3, 1, 34, 8), (128, 128, 3, 1, 18, 8), (384, 128, 3, 1, 18, 8), (256, 256, 3, 1, 10, 8), (256, 512, 4, 1, 19, 8),
          (16, 32, 3, 2, 130, 8), (64, 128, 4, 2, 66, 8), (128, 256, 4, 2, 34, 8), (48, 16, 1, 1, 128, 8), (1, 16, 3, 1, 130, 8), (1, 64, 4, 2, 130, 8)]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for (ci, co, k, s, S, N) in shapes:
    if only and only != "%d-%d" % (ci, co):
        continue
    if mode == "dgrad" and ci == 1:
        continue      # the network input needs no gradient in this benchmark: nothing to time
    net = E.Network("t", OrderedDict([("c.w", (k, k, k, ci, co)), ("c.b", (co,))]))
    net.load({"c.w": np.random.default_rng(0).standard_normal((k, k, k, ci, co)).astype(np.float32) * 0.05, "c.b": np.zeros(co, np.float32)})
    layer = E.Conv3D(net, "c", k, s, ci, co)
    net.repack()
    x = torch.randn((N, S, S, S, ci), device="cuda").to(torch.bfloat16 if ci > 1 else torch.float32)
    O = (S - k) // s + 1
    y = torch.empty((N, O, O, O, co), device="cuda", dtype=torch.bfloat16)
    dy = torch.randn_like(y)
    dx = torch.empty_like(x)
    desc = layer.desc(N, S, S, S)
    flops = 2.0 * k ** 3 * ci * co * N * O ** 3
    def run():
        if mode == "fwd":
            call("vg_conv3d_fwd", desc, x, layer.wf if ci > 1 else layer.w.w, layer.b.w, y)
        elif mode == "dgrad":
            call("vg_conv3d_dgrad", desc, dy, layer.wd, dx) if ci > 1 else None
        else:
            call("vg_conv3d_wgrad", desc, x, dy, layer.w.grad, layer.b.grad if ci == 1 else None)
    for _ in range(2):
        run()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    byt = (x.numel() + y.numel()) * 2
    print("%s %4d->%4d k%d s%d S=%3d N=%d : %8.3f ms  %7.1f TFLOP/s   (min-bytes %.1f GB/s)" % (mode, ci, co, k, s, S, N, t, flops / t / 1e9, byt / t / 1e6), flush=True)
