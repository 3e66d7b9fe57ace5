import sys, os, torch, numpy as np
from collections import OrderedDict
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nets as ON
from van_gan_b200.vnet_model import custom_vnet
from van_gan_b200 import engine as E
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b))
for (S, filters, L, N) in [(32, 16, 2, 1), (32, 16, 3, 1), (32, 16, 4, 1), (64, 16, 4, 1)]:
    rng = np.random.default_rng(1)
    shapes = ON.vnet_param_shapes(filters, L, 1)
    init = ON.init_params(shapes, 5, 0.05)
    x = torch.tensor(rng.standard_normal((N, S, S, S, 1)), dtype=torch.float32).clamp(-1, 1)
    masks = ON.make_vnet_masks(rng, N, filters, L)
    gy = torch.tensor(rng.standard_normal((N, S, S, S, 1)), dtype=torch.float32)
    res = {}
    for emu in (False, True):
        ON.Emu.on = emu
        P = ON.to_torch(init)
        y = ON.vnet_forward(P, x, L, masks)
        g = torch.autograd.grad((y * gy).sum(), list(P.values()))
        ON.Emu.on = False
        res[emu] = (y.detach(), OrderedDict(zip(P.keys(), g)))
    net = custom_vnet((S, S, S, 1), use_batch_norm=False, upsample_mode='upsample', dropout=0.5, filters=filters, num_layers=L, output_activation='tanh')
    net.load(init)
    tape = E.Tape()
    out = net.forward(tape, E.Var(x.cuda()), training=True, masks=[m.cuda() for m in masks])
    net.zero_grad()
    tape.backward([(out, gy.cuda())], net.trainable_variables)
    gg = net.export_grads()
    for emu in (False, True):
        y, g = res[emu]
        num = sum(float(((torch.tensor(gg[k]).double() - g[k].double()) ** 2).sum()) for k in shapes)
        den = sum(float((g[k].double() ** 2).sum()) for k in shapes)
        worst = max((rel(torch.tensor(gg[k]), g[k]), k) for k in shapes)
        print("S=%d L=%d vs %s: out %.4g  grads(all) %.4g  worst %s" % (S, L, "emu" if emu else "fp32", rel(out.data, y), (num / den) ** 0.5, worst))
    print("   emu vs fp32 oracle: out %.4g" % rel(res[True][0], res[False][0]), flush=True)
