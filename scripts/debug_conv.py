import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from collections import OrderedDict
from test_gpu_kernels import CONV_CASES, rel_l2, _bf
from oracle import nets as ON
from van_gan_b200 import engine as E
from van_gan_b200._lib import ACT_NONE, ACT_TANH
for (Cin, Cout, K, stride, sp, N) in CONV_CASES:
    rng = np.random.default_rng(Cin * 1000 + Cout + K)
    x = torch.tensor(rng.standard_normal((N,) + sp + (Cin,)), dtype=torch.float32)
    w = torch.tensor(ON.he_normal(rng, (K, K, K, Cin, Cout)), dtype=torch.float32)
    b = torch.tensor(0.1 * rng.standard_normal(Cout), dtype=torch.float32)
    xd = x if Cin == 1 else _bf(x)
    wd = w if (Cin == 1) else _bf(w)
    xr, wr, br = xd.clone().requires_grad_(True), wd.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = ON.conv3d(xr, wr, br, stride=stride)
    gy = torch.tensor(rng.standard_normal(y.shape), dtype=torch.float32)
    if Cout != 1:
        gy = _bf(gy)
    y.backward(gy)
    net = E.Network("t", OrderedDict([("c.w", tuple(w.shape)), ("c.b", (Cout,))]))
    net.load({"c.w": wd.numpy(), "c.b": b.numpy()})
    layer = E.Conv3D(net, "c", K, stride, Cin, Cout)
    net.repack()
    tape = E.Tape()
    xv = E.Var(x.cuda() if Cin == 1 else x.to(torch.bfloat16).cuda())
    out = layer(tape, xv)
    yref = y.detach() if Cout == 1 else _bf(y.detach())
    tape.backward([(out, gy.cuda() if Cout == 1 else gy.to(torch.bfloat16).cuda())], net.trainable_variables, wrt_vars=[xv])
    dxref = xr.grad if Cin == 1 else _bf(xr.grad)
    print("%4d->%4d k%d s%d  fwd %.2e  dgrad %.2e  wgrad %.2e  bgrad %.2e" % (Cin, Cout, K, stride, rel_l2(out.data.float(), yref),
          rel_l2(xv.grad.float(), dxref), rel_l2(net.params["c.w"].grad, wr.grad), rel_l2(net.params["c.b"].grad, br.grad)))
