"""Teacher-forced block decomposition of the ResUNet generator and the PatchGAN discriminator.

TEST INFRASTRUCTURE.  The whole-network gradient of a random-init, ~58-stage-deep ResUNet is ill-conditioned
with respect to ANY bf16 rounding (bf16 weight operands alone move the stem gradients by
~18 % against the fp64 oracle), so a whole-network 2e-2 bound cannot separate "wiring error" from "bf16".
Per block it can: every block below is 1-3 convolutions deep, is fed the ORACLE's input activation and the same
upstream gradient on both sides, and its input gradient and parameter gradients are compared with the fp32
oracle at north_star's relative-L2 2e-2.

`oracle_blocks(P)`  -> ordered {block name: (fn(P, *inputs) -> output, input names, parameter prefixes)}
The CUDA-side runners live in the GPU test (they need the extension); both sides use the same block list.

Follows resunet_model.py:69-100 (stem), :103-143 (residual_block), :146-182 (upsample_concat_block),
:236-238 (bridge), :245 (head); discriminator.py:47-124; building_blocks.py:126-196.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

from oracle import nets as ON

NUM_LAYERS = 4


# ----------------------------------------------------------------------------- ResUNet blocks (oracle side)
def gen_stem(p, x):
    conv = ON.conv3d(ON.reflect_pad(x), p["stem.conv0.w"], p["stem.conv0.b"])
    conv = ON._conv_block(p, "stem.cb", conv)
    sc = ON.conv3d(x, p["stem.short.conv.w"], p["stem.short.conv.b"], padding="same")
    sc = ON._norm_act(p, "stem.short.in", sc, act=False)
    return ON._qa(conv + sc)


def gen_enc(e):
    return lambda p, x: ON._res_block(p, "enc%d" % e, x, 2)


def gen_bridge(p, x):
    return ON._conv_block(p, "bridge2", ON._conv_block(p, "bridge1", x))


def gen_dec(d):
    return lambda p, lo, skip: ON._res_block(p, "dec%d" % d, torch.cat([ON.upsample2(lo), skip], dim=-1), 1)


def gen_head(p, x):
    return torch.tanh(ON.conv3d(x, p["head.w"], p["head.b"], padding="same"))


def gen_blocks():
    """name -> (oracle fn, names of the taps that feed it, parameter-name prefixes it owns)"""
    B = OrderedDict()
    B["stem"] = (gen_stem, ["input"], ["stem."])
    for e in range(1, NUM_LAYERS + 1):
        B["enc%d" % e] = (gen_enc(e), ["stem" if e == 1 else "enc%d" % (e - 1)], ["enc%d." % e])
    B["bridge"] = (gen_bridge, ["enc%d" % NUM_LAYERS], ["bridge1.", "bridge2."])
    for d in reversed(range(NUM_LAYERS)):
        lo = "bridge" if d == NUM_LAYERS - 1 else "dec%d" % (d + 1)
        B["dec%d" % d] = (gen_dec(d), [lo, "stem" if d == 0 else "enc%d" % d], ["dec%d." % d])
    B["head"] = (gen_head, ["dec0"], ["head."])
    return B


# ----------------------------------------------------------------------------- PatchGAN stages (oracle side)
# The CUDA path fuses InstanceNorm + LeakyReLU + SpatialDropout3D + padding + GaussianNoise of stage k-1 into the producer of
# conv k's input, so a stage here runs from one raw convolution output to the next.
def disc_stage(k):
    def f(p, h, noise, masks):
        if k == 0:
            h = ON.reflect_pad(h) + noise[0]                                               # discriminator.py:50-52
            return ON.conv3d(h, p["d0.conv.w"], p["d0.conv.b"], stride=2)                  # :63-69
        h = F.leaky_relu(ON.instance_norm(h, p["d%d.in.gamma" % (k - 1)], p["d%d.in.beta" % (k - 1)]), 0.2)
        if k >= 2:
            h = h * masks[k - 2]                                                           # SpatialDropout3D of block k-1
        if k <= 2:
            h = ON._qa(ON.reflect_pad(h) + noise[k])                                       # building_blocks.py:165-170
            return ON.conv3d(h, p["d%d.conv.w" % k], None, stride=2)
        h = ON._qa(h + noise[k])
        if k == 3:
            return ON.conv3d(h, p["d3.conv.w"], None, stride=1, padding="same")            # discriminator.py:91-103
        return ON.conv3d(h, p["dout.conv.w"], p["dout.conv.b"], padding="same")            # :108-114
    return f


def disc_stage_params(k):
    names = []
    if k >= 1:
        names += ["d%d.in.gamma" % (k - 1), "d%d.in.beta" % (k - 1)]
    names += ["d%d.conv.w" % k] if k < 4 else ["dout.conv.w", "dout.conv.b"]
    if k == 0:
        names += ["d0.conv.b"]
    return names


def disc_stage_inputs(p, x, noise, masks):
    """Raw convolution outputs that feed stages 1..4 (fp32 oracle), plus the input for stage 0."""
    ins = [x]
    h = x
    for k in range(4):
        h = disc_stage(k)(p, h, noise, masks)
        ins.append(h.detach())
    return ins


# ----------------------------------------------------------------------------- helpers shared by both sides
def bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


def smooth_grad(rng, shape, scale=1.0):
    """Upstream gradient with both a smooth and a white component (bf16-exact values)."""
    from scipy import ndimage
    import numpy as np
    g = rng.standard_normal(shape)
    g = 0.5 * g + ndimage.gaussian_filter(g, (0, 1, 1, 1, 0)) * 3.0
    return bf(torch.tensor(scale * g, dtype=torch.float32)) if shape[-1] > 1 else torch.tensor(scale * g, dtype=torch.float32)


def agg_rel(gk, go, names=None):
    names = list(go.keys()) if names is None else names
    num = sum(float(((torch.as_tensor(gk[n]).double().cpu() - torch.as_tensor(go[n]).double()) ** 2).sum()) for n in names)
    den = sum(float((torch.as_tensor(go[n]).double() ** 2).sum()) for n in names)
    return (num / max(den, 1e-300)) ** 0.5


def cosine(gk, go, names=None):
    names = list(go.keys()) if names is None else names
    dot = sum(float((torch.as_tensor(gk[n]).double().cpu() * torch.as_tensor(go[n]).double()).sum()) for n in names)
    a = sum(float((torch.as_tensor(gk[n]).double() ** 2).sum()) for n in names)
    b = sum(float((torch.as_tensor(go[n]).double() ** 2).sum()) for n in names)
    return dot / max((a * b) ** 0.5, 1e-300)


def oracle_block_grads(fn, p, inputs, gout, pnames):
    """Runs one block on detached inputs; returns (output, {input i: grad}, {param: grad})."""
    xs = [t.clone().requires_grad_(t.shape[-1] > 1 or True) for t in inputs]
    y = fn(p, *xs)
    params = [p[n] for n in pnames]
    g = torch.autograd.grad(y, xs + params, gout, allow_unused=True)
    gx = [gi if gi is not None else torch.zeros_like(x) for gi, x in zip(g[:len(xs)], xs)]
    gp = {n: (gi if gi is not None else torch.zeros_like(p[n])) for n, gi in zip(pnames, g[len(xs):])}
    return y.detach(), gx, gp


K_PROJ = 64


def projections(flat, seed, k=K_PROJ):
    """k dot products of `flat` (1-D float64 CPU tensor) with seeded +-1 vectors r_i.  E[(r.a)(r.b)] = a.b, so from the projections of
    two gradients one estimates ||a - b||^2 = mean_i (r_i.a - r_i.b)^2 and a.b = mean_i (r_i.a)(r_i.b) (relative std ~ sqrt(2/k))."""
    import numpy as np
    out = np.zeros(k)
    g = torch.Generator()
    g.manual_seed(seed)
    for i in range(k):
        signs = torch.randint(0, 2, (flat.numel(),), generator=g, dtype=torch.int8).to(torch.float64) * 2 - 1
        out[i] = float((signs * flat).sum())
    return out
