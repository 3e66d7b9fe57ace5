"""How much does bf16 move the GRADIENTS of this network, independent of any CUDA code?  (CPU only, oracle only.)

Arbiter: the fp64 oracle.  Prints, for one ResUNet application at S^3 and for one residual block,
  * fp32 vs fp64                                   (the reference's own arithmetic noise)
  * bf16 weight operands only vs fp64              ('weights')
  * bf16 at every storage point of the CUDA path   ('both': oracle.nets.Emu)
  * forward-only / gradient-only rounding
Result at 64^3: one conv block 0.2-0.3 %, two chained conv blocks 3-4.5 % (ReLU sign flips: the
gradient error is ~sqrt(fraction of units whose pre-activation changes sign)), whole network 18 % (weights only) to 28 % (all
storage points), whole train step 36-51 %.  TEST INFRASTRUCTURE.

    python scripts/grad_conditioning.py [S=64]
"""
import os
import sys

import numpy as np
import torch
from scipy import ndimage

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import nets as ON  # noqa: E402
import _blocks as B  # noqa: E402
from test_gpu_train_step import synth  # noqa: E402


class RoundFwdOnly(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


class RoundBwdOnly(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.clone()

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


class NoRound(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.clone()

    @staticmethod
    def backward(ctx, g):
        return g


ORIG = ON._RoundBoth
MODES = {"both": ORIG, "fwd": RoundFwdOnly, "bwd": RoundBwdOnly, "weights": NoRound}


def with_mode(mode, fn):
    ON.Emu.on = mode != "off"
    ON._RoundBoth = MODES.get(mode, ORIG)
    try:
        return fn()
    finally:
        ON.Emu.on = False
        ON._RoundBoth = ORIG


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    rng = np.random.default_rng(3)
    real_I, _ = synth(rng, 1, S)
    init = ON.init_params(ON.resunet_param_shapes(), 1, 0.05)
    g_up = torch.tensor(ndimage.gaussian_filter(rng.standard_normal((1, S, S, S, 1)), (0, 1, 1, 1, 0)), dtype=torch.float32)

    def whole(dtype):
        P = ON.to_torch(init, dtype=dtype)
        y = ON.resunet_forward(P, real_I.to(dtype))
        return y.detach(), dict(zip(P.keys(), torch.autograd.grad(y, list(P.values()), g_up.to(dtype))))

    y64, g64 = whole(torch.float64)
    print("one ResUNet application at %d^3, gradient w.r.t. all parameters, against the fp64 oracle" % S)
    for mode in ("off", "weights", "fwd", "bwd", "both"):
        y, g = with_mode(mode, lambda: whole(torch.float32))
        print("  %-8s forward rel-L2 %.2e   gradient rel-L2 %.4f   cosine %.4f" % (
            mode, float((y.double() - y64).norm() / y64.norm()), B.agg_rel(g, g64), B.cosine(g, g64)))

    P = ON.to_torch(init)
    taps = {}
    with torch.no_grad():
        ON.resunet_forward(P, real_I, taps=taps)
    x = B.bf(taps["enc1"])
    name = "enc2"
    cases = {
        "one conv block (enc2.cb1)": (lambda p, h: ON._conv_block(p, name + ".cb1", h, 2), name + ".cb1"),
        "two chained conv blocks": (lambda p, h: ON._conv_block(p, name + ".cb2", ON._conv_block(p, name + ".cb1", h, 2), 1), name + ".cb"),
        "residual block enc2": (lambda p, h: ON._res_block(p, name, h, 2), name + "."),
    }
    print("teacher-forced sub-networks (same input, same upstream gradient), against the fp32 oracle")
    for label, (fn, prefix) in cases.items():
        pn = [n for n in P if n.startswith(prefix)]
        with torch.no_grad():
            shape = tuple(fn(P, x).shape)
        g = B.smooth_grad(rng, shape)
        y, gx, gp = B.oracle_block_grads(fn, P, [x], g, pn)
        for mode in ("weights", "both"):
            ye, gxe, gpe = with_mode(mode, lambda: B.oracle_block_grads(fn, P, [x], g, pn))
            flips = float(((ye > 0) != (y > 0)).float().mean())
            print("  %-28s %-8s forward %.2e   dx %.4f   params %.4f   output sign flips %.2e" % (
                label, mode, float((ye - y).norm() / y.norm()), float((gxe[0] - gx[0]).norm() / gx[0].norm()), B.agg_rel(gpe, gp), flips))


if __name__ == "__main__":
    main()
