"""Isolated InstanceNorm kernel timing: achieved GB/s (algorithmic bytes) for stats / apply / bwd at the big generator shapes."""
import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from van_gan_b200 import _lib
from van_gan_b200._lib import call, InDesc, ACT_RELU, ACT_NONE, PAD_REFLECT, PAD_ZERO
L = _lib.lib()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
def timeit(fn, n=5):
    for _ in range(2): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
if len(sys.argv) > 1 and sys.argv[1] == 'one':
    shapes = [(8, 128, 16, 1)]
else:
  shapes = [(8, 128, 16, 1), (8, 64, 32, 1), (8, 128, 48, 1), (8, 32, 64, 1), (1, 128, 16, 1), (8, 128, 16, 0)]
for (N, S, C, pad) in shapes:
    x = torch.randn((N, S, S, S, C), device="cuda").to(torch.bfloat16)
    P = S + 2 * pad
    y = torch.empty((N, P, P, P, C), device="cuda", dtype=torch.bfloat16)
    dy = torch.randn_like(y)
    dx = torch.empty_like(x)
    res = torch.randn_like(x) if pad == 0 else None
    dres = torch.empty_like(x) if pad == 0 else None
    mean = torch.empty(N * C, device="cuda"); rstd = torch.empty(N * C, device="cuda")
    gamma = torch.ones(C, device="cuda"); beta = torch.zeros(C, device="cuda")
    dg = torch.zeros(C, device="cuda"); db = torch.zeros(C, device="cuda")
    wsb = L.vg_instnorm_workspace_bytes(N, S, S, S, C)
    ws = torch.empty(wsb // 4 + 1, device="cuda")
    desc = InDesc(N, S, S, S, C, _lib.VG_BF16, ACT_RELU if pad else ACT_NONE, 0.2, pad, pad, PAD_REFLECT, 0.0, 0)
    desc_b = InDesc(N, S, S, S, C, _lib.VG_BF16 | _lib.IN_DY_SCRATCH, ACT_RELU if pad else ACT_NONE, 0.2, pad, pad, PAD_REFLECT, 0.0, 0)
    el = x.numel(); elp = y.numel()
    t = timeit(lambda: call("vg_instnorm_stats", x, _lib.VG_BF16, N, S, S, S, C, mean, rstd, ws, wsb))
    print("stats N=%d S=%d C=%d: %.3f ms %.0f GB/s" % (N, S, C, t, el * 2 / t / 1e6))
    t = timeit(lambda: call("vg_instnorm_apply", desc, x, res, y, mean, rstd, gamma, beta, None, None))
    print("apply N=%d S=%d C=%d pad=%d: %.3f ms %.0f GB/s" % (N, S, C, pad, t, (el * 2 * (2 if res is not None else 1) + elp * 2) / t / 1e6))
    t = timeit(lambda: call("vg_instnorm_bwd", desc_b, dy, x, mean, rstd, gamma, beta, None, dx, 0, dres, dg, db, ws, wsb))
    byt = (el * 2 + elp * 2) * 2 + el * 2 * (2 if dres is not None else 1)
    print("bwd   N=%d S=%d C=%d pad=%d: %.3f ms %.0f GB/s" % (N, S, C, pad, t, byt / t / 1e6), flush=True)
