"""GPU parity of the full VAN-GAN train step (2 ResUNet generators + 2 PatchGAN discriminators,
all ten losses, four backward sweeps, clip+Adam) against the fp32 CPU oracle on identical inputs,
weights, discriminator noise and dropout masks.

Tolerance (north_star): losses and per-network gradients within relative L2 2e-2 (bf16 operands,
fp32 accumulation)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


class Args:
    def __init__(self, S, G, nd):
        self.N_DEVICES, self.GLOBAL_BATCH_SIZE = nd, G
        self.INPUT_IMG_SIZE = (G, S, S, S, 1)
        self.CHANNELS, self.DIMENSIONS = 1, 3
        self.SUBVOL_PATCH_SIZE = (S, S, S)
        self.train_steps, self.BATCH_SIZE, self.output_dir = 1, G // nd, "/tmp"


def synth(rng, n, S):
    """smooth noise in [-1,1] (imaging domain) and soft tubes in [-1,1] (segmentation domain)"""
    from scipy import ndimage
    I = np.stack([ndimage.gaussian_filter(rng.standard_normal((S, S, S)), 2.0) for _ in range(n)])
    I = 2 * (I - I.min(axis=(1, 2, 3), keepdims=True)) / np.ptp(I, axis=(1, 2, 3), keepdims=True) - 1
    zz, yy, xx = np.mgrid[0:S, 0:S, 0:S]
    Sg = []
    for _ in range(n):
        v = np.zeros((S, S, S))
        for _t in range(6):
            p0, d = rng.random(3) * S, rng.standard_normal(3)
            d /= np.linalg.norm(d)
            rel = np.stack([zz - p0[0], yy - p0[1], xx - p0[2]], -1)
            dist = np.linalg.norm(rel - (rel @ d)[..., None] * d, axis=-1)
            v = np.maximum(v, 1.0 / (1.0 + np.exp((dist - (1.5 + 2 * rng.random())) * 2.0)))
        Sg.append(2 * v - 1 + 1e-3 * rng.standard_normal(v.shape))
    return (torch.tensor(I[..., None], dtype=torch.float32), torch.tensor(np.stack(Sg)[..., None], dtype=torch.float32))


@pytest.mark.parametrize("S,b", [(32, 1), (32, 2)])
def test_train_step_matches_oracle(cuda, S, b):
    from oracle import losses as OL, nets as ON, step as OS
    from van_gan_b200.vangan import VanGan
    rng = np.random.default_rng(100 + b)
    nd = 2            # pretend to be one of two replicas: exercises the n_devices / global-batch scalings
    G = b * nd
    real_I, real_S = synth(rng, b, S)
    init = {"gen_IS": ON.init_params(ON.resunet_param_shapes(), 1, 0.05), "gen_SI": ON.init_params(ON.resunet_param_shapes(), 2, 0.05),
            "disc_I": ON.init_params(ON.disc_param_shapes(), 3, 0.05), "disc_S": ON.init_params(ON.disc_param_shapes(), 4, 0.05)}
    rand = {k: ON.make_disc_rand(rng, b, S) for k in ("S_real", "S_fake", "I_real", "I_fake")}

    cfg = OL.make_cfg(G, nd)
    P = {k: ON.to_torch(v) for k, v in init.items()}
    res_o, grads_o, aux_o = OS.replica_grads(cfg, P, real_I, real_S, rand)

    gan = VanGan(Args(S, G, nd), gen_i2s='resUnet', gen_s2i='resUnet')
    for k, net in gan.networks.items():
        net.load(init[k])
    rand_d = {k: ([t.cuda() for t in nz], [m.cuda() for m in mk]) for k, (nz, mk) in rand.items()}
    res_k = gan.train_step(real_I, real_S, rand=rand_d, apply=False)

    for k in OS.RESULT_KEYS:
        o = float(res_o[k])
        assert abs(res_k[k] - o) <= 2e-2 * abs(o) + 1e-4, (k, res_k[k], o)
    for name, net in gan.networks.items():
        g = net.export_grads()
        num = sum(float(((torch.tensor(g[n]).double() - grads_o[name][n].double()) ** 2).sum()) for n in g)
        den = sum(float((grads_o[name][n].double() ** 2).sum()) for n in g)
        rel = (num / den) ** 0.5
        assert rel < 2e-2, (name, rel)


def test_adam_update_and_second_step(cuda):
    """two consecutive optimizer steps stay within tolerance of the oracle's weights"""
    from oracle import losses as OL, nets as ON, step as OS
    from van_gan_b200.vangan import VanGan
    rng = np.random.default_rng(5)
    S, b = 32, 1
    real_I, real_S = synth(rng, b, S)
    init = {"gen_IS": ON.init_params(ON.resunet_param_shapes(), 1), "gen_SI": ON.init_params(ON.resunet_param_shapes(), 2),
            "disc_I": ON.init_params(ON.disc_param_shapes(), 3), "disc_S": ON.init_params(ON.disc_param_shapes(), 4)}
    cfg = OL.make_cfg(b, 1)
    P = {k: ON.to_torch(v) for k, v in init.items()}
    opts = {k: OS.Adam(list(v.keys())) for k, v in init.items()}
    gan = VanGan(Args(S, b, 1), gen_i2s='resUnet', gen_s2i='resUnet')
    for k, net in gan.networks.items():
        net.load(init[k])
    for it in range(2):
        rand = {k: ON.make_disc_rand(rng, b, S) for k in ("S_real", "S_fake", "I_real", "I_fake")}
        res_o, _ = OS.train_step_dp(cfg, P, opts, real_I, real_S, [rand])
        rand_d = {k: ([t.cuda() for t in nz], [m.cuda() for m in mk]) for k, (nz, mk) in rand.items()}
        res_k = gan.train_step(real_I, real_S, rand=rand_d)
        for k in OS.RESULT_KEYS:
            assert abs(res_k[k] - res_o[k]) <= 3e-2 * abs(res_o[k]) + 1e-3, (it, k, res_k[k], res_o[k])
    # Adam's first steps move every weight by ~lr regardless of gradient scale, so compare the
    # displacement direction: the update must correlate strongly with the oracle's
    for name, net in gan.networks.items():
        w = net.export()
        num = den1 = den2 = 0.0
        for n in w:
            dk = torch.tensor(w[n] - init[name][n]).double().flatten()
            do = (P[name][n].detach() - torch.tensor(init[name][n])).double().flatten()
            num += float(dk @ do); den1 += float(dk @ dk); den2 += float(do @ do)
        assert num / (den1 * den2) ** 0.5 > 0.9, (name, num / (den1 * den2) ** 0.5)
