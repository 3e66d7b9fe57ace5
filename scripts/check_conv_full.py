"""Full-size conv parity against torch (cuDNN fp32 on the same bf16-rounded operands): exercises persistent CTAs with many work
items each, which the small unit-test shapes do not.  Prints rel-L2 of fwd and dgrad per shape."""
import sys, os, torch, numpy as np
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collections import OrderedDict
from van_gan_b200 import engine as E
from van_gan_b200._lib import call
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
shapes = [(16, 16, 3, 1, 130, 2), (48, 16, 3, 1, 130, 2), (32, 32, 3, 1, 66, 8), (96, 32, 3, 1, 66, 8), (16, 32, 3, 2, 130, 2),
          (64, 128, 4, 2, 66, 4), (128, 256, 4, 2, 34, 8), (256, 512, 4, 1, 19, 8), (64, 64, 3, 1, 34, 8), (48, 16, 1, 1, 128, 2)]
def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
worst = 0.0
for (ci, co, k, s, S, N) in shapes:
    g = torch.Generator(device="cuda").manual_seed(ci * 7 + co)
    w = (torch.randn((k, k, k, ci, co), device="cuda", generator=g) * (2.0 / (k ** 3 * ci)) ** 0.5).to(torch.bfloat16).float()
    net = E.Network("t", OrderedDict([("c.w", (k, k, k, ci, co)), ("c.b", (co,))]))
    net.load({"c.w": w.cpu().numpy(), "c.b": np.zeros(co, np.float32)})
    layer = E.Conv3D(net, "c", k, s, ci, co)
    net.repack()
    x = torch.randn((N, S, S, S, ci), device="cuda", generator=g).to(torch.bfloat16)
    O = (S - k) // s + 1
    y = torch.empty((N, O, O, O, co), device="cuda", dtype=torch.bfloat16)
    dy = torch.randn((N, O, O, O, co), device="cuda", generator=g).to(torch.bfloat16)
    dx = torch.empty_like(x)
    desc = layer.desc(N, S, S, S)
    call("vg_conv3d_fwd", desc, x, layer.wf, layer.b.w, y)
    call("vg_conv3d_dgrad", desc, dy, layer.wd, dx)
    wt = w.permute(4, 3, 0, 1, 2).contiguous()                      # (co, ci, kd, kh, kw)
    rf = rd = 0.0
    for n in range(N):                                              # per sample: bounds the fp32 reference's memory
        xn = x[n:n + 1].float().permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)
        yr = F.conv3d(xn, wt, stride=s)
        yr.backward(dy[n:n + 1].float().permute(0, 4, 1, 2, 3).contiguous())
        rf = max(rf, rel(y[n].float(), yr[0].permute(1, 2, 3, 0).to(torch.bfloat16).float()))
        rd = max(rd, rel(dx[n].float(), xn.grad[0].permute(1, 2, 3, 0).to(torch.bfloat16).float()))
        del xn, yr
    worst = max(worst, rf, rd)
    print("%4d->%4d k%d s%d S=%3d N=%d  fwd rel-L2 %.2e  dgrad rel-L2 %.2e" % (ci, co, k, s, S, N, rf, rd), flush=True)
print("WORST %.2e %s" % (worst, "OK" if worst < 2e-3 else "FAIL"))
