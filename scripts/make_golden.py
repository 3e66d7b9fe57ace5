"""Generates tests/golden/*.npz from the CPU oracle (seeded).  The reference ships no golden vectors
(and TensorFlow cannot run here), so these freeze the ORACLE's outputs: they pin the oracle against
accidental edits and give the CUDA path fixed vectors that travel to the GPU box.
usage: python scripts/make_golden.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import losses as OL, nets as ON, step as OS, np_ref

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(1)          # fixed reduction order

# 1. soft skeleton + clDice/cycle losses on a 20^3 pair
rng = np.random.default_rng(42)
S = 20
real = (rng.random((1, S, S, S, 1)) * 2 - 1).astype(np.float32)
cyc = np.tanh(rng.standard_normal((1, S, S, S, 1))).astype(np.float32)
x01 = rng.random((1, S, S, S, 1)).astype(np.float32)
cfg = OL.make_cfg(2, 2)
c = torch.tensor(cyc, requires_grad=True)
r = torch.tensor(real)
vals, grads = {}, {}
for name, fn in (("bce", lambda: OL.cycle_loss(cfg, r, c, typ="bce")), ("mse", lambda: OL.cycle_loss(cfg, r, c, typ="mse")),
                 ("ssim", lambda: OL.cycle_reconstruction(cfg, r, c)), ("seg", lambda: OL.cycle_seg_loss(cfg, r, c, iters=5))):
    c.grad = None
    l = fn()
    l.backward()
    vals[name] = np.float64(l.item())
    grads[name] = c.grad.numpy().copy()
np.savez_compressed(os.path.join(OUT, "losses_20.npz"), real=real, cycled=cyc, x01=x01,
                    skel5=OL.soft_skel(torch.tensor(x01), 5).numpy(), erode=OL.soft_erode(torch.tensor(x01)).numpy(),
                    **{"val_" + k: v for k, v in vals.items()}, **{"grad_" + k: v for k, v in grads.items()})

# 2. one ResUNet application and one discriminator application at 16^3 / 32^3 (weights from seeds)
S = 32
rng = np.random.default_rng(43)
xg = np.clip(rng.standard_normal((1, S, S, S, 1)), -1, 1).astype(np.float32)
Pg = ON.to_torch(ON.init_params(ON.resunet_param_shapes(), 7, 0.05), requires_grad=False)
yg = ON.resunet_forward(Pg, torch.tensor(xg)).numpy()
Pd = ON.to_torch(ON.init_params(ON.disc_param_shapes(), 8, 0.05), requires_grad=False)
nz, mk = ON.make_disc_rand(rng, 1, S)
yd = ON.disc_forward(Pd, torch.tensor(xg), nz, mk).numpy()
yd_inf = ON.disc_forward(Pd, torch.tensor(xg)).numpy()
np.savez_compressed(os.path.join(OUT, "nets_32.npz"), x=xg, gen_out=yg, disc_out=yd, disc_out_inference=yd_inf,
                    **{"noise%d" % i: t.numpy() for i, t in enumerate(nz)}, **{"mask%d" % i: t.numpy() for i, t in enumerate(mk)})

# 3. stitching of a 40x36x24 volume with a cheap analytic "generator"
rng = np.random.default_rng(44)
vol = rng.random((40, 36, 24, 1)).astype(np.float32)
gen = lambda a: np.tanh(1.5 * a - 0.3)
a = np_ref.stitch_subvolumes(gen, vol, (1, 16, 16, 16, 1), stride=(8, 8, 8), complete=True, padFactor=0.25)
b = np_ref.stitch_subvolumes(gen, vol, (1, 16, 16, 16, 1), stride=(8, 8, 8), complete=False)
np.savez_compressed(os.path.join(OUT, "stitch_40.npz"), vol=vol, complete=a, plain=b)
print("wrote", os.listdir(OUT))
