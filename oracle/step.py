"""Oracle: VanGan.compute_losses / train_step / data-parallel emulation (torch CPU autograd).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows vangan.py:270-353 (compute_losses), :380-440 (train_step, non-Wasserstein branch
:425-438: four `minimize` calls, each differentiating ONE loss w.r.t. ONE network's variables on a
persistent tape), :459-490 (replica SUM of the result dict) and the Keras OptimizerV2 Adam that
`minimize` drives (vangan.py:220-235: lr 2e-4, beta_1 0.5, beta_2 0.9, clipnorm 100).
"""
from collections import OrderedDict

import numpy as np
import torch

from . import losses as L
from . import nets

RESULT_KEYS = ("total_IS_loss", "total_SI_loss", "D_I_loss", "D_S_loss", "gen_IS_loss", "gen_SI_loss",
               "cycle_gen_SIS_loss", "cycle_gen_ISI_loss", "seg_loss", "reconstruction_loss_I")


def compute_losses(cfg, P, real_I, real_S, rand=None, iters=15):
    """P: dict with 'gen_IS','gen_SI','disc_I','disc_S' -> param dicts (torch tensors).
    rand: None (inference mode) or dict with keys 'S_real','S_fake','I_real','I_fake' ->
    (noise list, mask list) for the four discriminator applications (vangan.py:315-319)."""
    def D(net, x, key):
        nz, mk = (None, None) if rand is None else rand[key]
        return nets.disc_forward(P[net], x, nz, mk)

    fake_S = nets.resunet_forward(P["gen_IS"], real_I)
    fake_I = nets.resunet_forward(P["gen_SI"], real_S)
    cycled_S = nets.resunet_forward(P["gen_IS"], fake_I)
    cycle_loss_I = L.cycle_loss(cfg, real_S, cycled_S, typ="bce")
    seg_loss = L.cycle_seg_loss(cfg, real_S, cycled_S, iters=iters)
    cycled_I = nets.resunet_forward(P["gen_SI"], fake_S)
    cycle_loss_S = L.cycle_loss(cfg, real_I, cycled_I, typ="mse")
    recon = L.cycle_reconstruction(cfg, real_I, cycled_I)

    disc_real_S = D("disc_S", real_S, "S_real")
    disc_fake_S = D("disc_S", fake_S, "S_fake")
    disc_real_I = D("disc_I", real_I, "I_real")
    disc_fake_I = D("disc_I", fake_I, "I_fake")

    gen_IS_loss = L.generator_loss_fn(cfg, disc_fake_S)
    gen_SI_loss = L.generator_loss_fn(cfg, disc_fake_I)
    disc_I_loss = L.discriminator_loss_fn(cfg, disc_real_I, disc_fake_I)
    disc_S_loss = L.discriminator_loss_fn(cfg, disc_real_S, disc_fake_S)
    total_I = gen_IS_loss + cycle_loss_I + seg_loss
    total_S = gen_SI_loss + cycle_loss_S + recon
    result = OrderedDict([
        ("total_IS_loss", total_I), ("total_SI_loss", total_S), ("D_I_loss", disc_I_loss),
        ("D_S_loss", disc_S_loss), ("gen_IS_loss", gen_IS_loss), ("gen_SI_loss", gen_SI_loss),
        ("cycle_gen_SIS_loss", cycle_loss_I), ("cycle_gen_ISI_loss", cycle_loss_S),
        ("seg_loss", seg_loss), ("reconstruction_loss_I", recon)])
    aux = dict(fake_S=fake_S, fake_I=fake_I, cycled_S=cycled_S, cycled_I=cycled_I,
               disc_real_S=disc_real_S, disc_fake_S=disc_fake_S, disc_real_I=disc_real_I,
               disc_fake_I=disc_fake_I)
    return result, aux


def replica_grads(cfg, P, real_I, real_S, rand, iters=15):
    """One replica's losses and the four per-network gradient dicts (before any all-reduce)."""
    result, aux = compute_losses(cfg, P, real_I, real_S, rand, iters)
    pairs = (("gen_IS", "total_IS_loss"), ("gen_SI", "total_SI_loss"),
             ("disc_I", "D_I_loss"), ("disc_S", "D_S_loss"))
    grads = {}
    for net, key in pairs:
        names = list(P[net].keys())
        g = torch.autograd.grad(result[key], [P[net][n] for n in names], retain_graph=True, allow_unused=True)
        grads[net] = OrderedDict((n, (torch.zeros_like(P[net][n]) if gi is None else gi)) for n, gi in zip(names, g))
    return result, grads, aux


def clip_by_norm(g, clip):
    """tf.clip_by_norm: g * clip / max(||g||, clip)."""
    n = torch.linalg.vector_norm(g.double()).to(g.dtype)
    return g * (clip / torch.maximum(n, torch.tensor(clip, dtype=g.dtype)))


class Adam:
    """Keras OptimizerV2 Adam (non-amsgrad), with per-variable clipnorm applied after aggregation."""

    def __init__(self, names, lr=2e-4, beta_1=0.5, beta_2=0.9, eps=1e-7, clipnorm=100.0):
        self.lr, self.b1, self.b2, self.eps, self.clip = lr, beta_1, beta_2, eps, clipnorm
        self.t = 0
        self.m = {n: None for n in names}
        self.v = {n: None for n in names}

    def apply(self, params, grads):
        self.t += 1
        lr_t = self.lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        with torch.no_grad():
            for n, w in params.items():
                g = clip_by_norm(grads[n], self.clip)
                if self.m[n] is None:
                    self.m[n] = torch.zeros_like(w)
                    self.v[n] = torch.zeros_like(w)
                self.m[n].mul_(self.b1).add_(g, alpha=1 - self.b1)
                self.v[n].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                w.sub_(lr_t * self.m[n] / (self.v[n].sqrt() + self.eps))


def train_step_dp(cfg, P, opts, real_I, real_S, rand_per_replica, iters=15):
    """Emulates MirroredStrategy: shard the global batch over cfg.n_devices replicas, run the
    replica step on each shard, SUM gradients and result entries (vangan.py:471-473), then one
    clip+Adam per network.  Returns (summed result dict of floats, summed grads)."""
    n = cfg.n_devices
    b = real_I.shape[0] // n
    tot_res, tot_g = None, None
    for r in range(n):
        sl = slice(r * b, (r + 1) * b)
        res, g, _ = replica_grads(cfg, P, real_I[sl], real_S[sl], rand_per_replica[r], iters)
        res = OrderedDict((k, float(v.detach())) for k, v in res.items())
        if tot_res is None:
            tot_res, tot_g = res, g
        else:
            for k in res:
                tot_res[k] += res[k]
            for net in g:
                for nm in g[net]:
                    tot_g[net][nm] = tot_g[net][nm] + g[net][nm]
    if opts is not None:
        for net in ("gen_IS", "gen_SI", "disc_I", "disc_S"):
            opts[net].apply(P[net], tot_g[net])
    return tot_res, tot_g
