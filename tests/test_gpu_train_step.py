"""GPU parity of the full VAN-GAN train step (2 ResUNet generators + 2 PatchGAN discriminators, all
ten losses, four backward sweeps, clip+Adam) against the fp32 CPU oracle on identical inputs, weights,
discriminator noise and dropout masks.

Tolerances (written here as north_star asks):
* the ten LOSSES: relative 2e-2 against the fp32 oracle (measured <= 9e-3).
* network FORWARD outputs: relative L2 5e-2 per generator application, 2e-2 per discriminator
  application (measured 3.0e-2 / 0.7e-2).  A ResUNet application is ~58 bf16 rounding stages deep
  and a random-init residual net amplifies relative perturbations ~linearly with depth.
* GRADIENTS: the conv / norm / loss kernels are individually exact to rounding
  (tests/test_gpu_kernels.py: 1e-3 .. 1e-5), but the full-step gradient is ill-conditioned w.r.t. ANY
  forward perturbation: ReLU/LeakyReLU derivative flips, and above all the reference's per-sample
  min-max normalisation, whose gradient puts a spike of ~1e3x the typical magnitude on the arg-min /
  arg-max voxel of the generated volume; a 1e-3 change of the forward moves that voxel.  The fp32
  reference run with TF32 tensor cores (TensorFlow's default on the GPUs it targets) is exposed to the
  same effect.  The criterion used is therefore relative to the noise floor of bf16 STORAGE itself:
  the oracle is re-run with bit-faithful bf16 rounding at exactly the points where the CUDA path
  stores bf16 (oracle.nets.Emu), and the CUDA gradients must be as close to the fp32 oracle as that
  emulation is (factor 1.5 + 2e-2 absolute slack), per network.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def agg_rel(gk, go):
    num = sum(float(((torch.as_tensor(gk[n]).double() - torch.as_tensor(go[n]).double()) ** 2).sum()) for n in gk)
    den = sum(float((torch.as_tensor(go[n]).double() ** 2).sum()) for n in gk)
    return (num / den) ** 0.5


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


def _setup(S, b, nd, seed, perturb=0.05):
    from oracle import nets as ON
    rng = np.random.default_rng(seed)
    real_I, real_S = synth(rng, b, S)
    init = {"gen_IS": ON.init_params(ON.resunet_param_shapes(), 1, perturb), "gen_SI": ON.init_params(ON.resunet_param_shapes(), 2, perturb),
            "disc_I": ON.init_params(ON.disc_param_shapes(), 3, perturb), "disc_S": ON.init_params(ON.disc_param_shapes(), 4, perturb)}
    rand = {k: ON.make_disc_rand(rng, b, S) for k in ("S_real", "S_fake", "I_real", "I_fake")}
    return real_I, real_S, init, rand


def _cuda_step(S, G, nd, init, real_I, real_S, rand, apply=False):
    from van_gan_b200.vangan import VanGan
    gan = VanGan(Args(S, G, nd), gen_i2s='resUnet', gen_s2i='resUnet')
    gan.keep_last = True
    for k, net in gan.networks.items():
        net.load(init[k])
    rand_d = {k: ([t.cuda() for t in nz], [m.cuda() for m in mk]) for k, (nz, mk) in rand.items()}
    res = gan.train_step(real_I, real_S, rand=rand_d, apply=apply)
    return gan, res


@pytest.mark.parametrize("S,b", [(32, 1), (32, 2), (64, 1)])
def test_train_step_matches_oracle(cuda, S, b):
    from oracle import losses as OL, nets as ON, step as OS
    nd = 2            # act as one of two replicas: exercises the n_devices / global-batch scalings
    G = b * nd
    real_I, real_S, init, rand = _setup(S, b, nd, 100 + b)
    cfg = OL.make_cfg(G, nd)
    res_o, grads_o, aux_o = OS.replica_grads(cfg, {k: ON.to_torch(v) for k, v in init.items()}, real_I, real_S, rand)
    ON.Emu.on = True
    try:
        res_e, grads_e, _ = OS.replica_grads(cfg, {k: ON.to_torch(v) for k, v in init.items()}, real_I, real_S, rand)
    finally:
        ON.Emu.on = False
    gan, res_k = _cuda_step(S, G, nd, init, real_I, real_S, rand)

    for k in OS.RESULT_KEYS:                                   # losses: 2e-2 vs the fp32 oracle
        o = float(res_o[k].detach())
        assert abs(res_k[k] - o) <= 2e-2 * abs(o) + 1e-4, (k, res_k[k], o)
    ftol = 5e-2 if S >= 64 else 8e-2      # at 32^3 the deepest level normalises over 2^3 voxels: ill-conditioned
    assert rel_l2(gan.last["fake_S"].data.cpu(), aux_o["fake_S"].detach()) < ftol
    assert rel_l2(gan.last["fake_I"].data.cpu(), aux_o["fake_I"].detach()) < ftol
    assert rel_l2(gan.last["disc_real_S"].data.cpu(), aux_o["disc_real_S"].detach()) < 2e-2
    assert rel_l2(gan.last["disc_real_I"].data.cpu(), aux_o["disc_real_I"].detach()) < 2e-2
    for name, net in gan.networks.items():                     # gradients: within the bf16-storage noise floor
        g = net.export_grads()
        floor = agg_rel(grads_e[name], grads_o[name])
        err = agg_rel(g, grads_o[name])
        assert err <= 1.5 * floor + 2e-2, (name, err, floor)


def test_single_application_gradients(cuda):
    """One generator / one discriminator application with a given upstream gradient: no chained
    applications and no min-max spikes, so the comparison is against the fp32 oracle directly."""
    from scipy import ndimage
    from oracle import nets as ON
    from van_gan_b200 import engine as E
    from van_gan_b200.discriminator import get_discriminator
    from van_gan_b200.resunet_model import ResUNet
    S = 64
    rng = np.random.default_rng(3)
    real_I, real_S = synth(rng, 1, S)
    init = ON.init_params(ON.resunet_param_shapes(), 1, 0.05)
    g_up = torch.tensor(ndimage.gaussian_filter(rng.standard_normal((1, S, S, S, 1)), (0, 1, 1, 1, 0)), dtype=torch.float32)

    def oracle_gen(emu):
        ON.Emu.on = emu
        try:
            P = ON.to_torch(init)
            y = ON.resunet_forward(P, real_I)
            return y.detach(), dict(zip(P.keys(), torch.autograd.grad(y, list(P.values()), g_up)))
        finally:
            ON.Emu.on = False

    yo, go = oracle_gen(False)
    _, ge = oracle_gen(True)
    net = ResUNet((S, S, S, 1), upsample_mode='simple')
    net.load(init)
    tape = E.Tape()
    out = net.forward(tape, E.Var(real_I.cuda()))
    net.zero_grad()
    tape.backward([(out, g_up.cuda())], net.trainable_variables)
    assert rel_l2(out.data.cpu(), yo) < 5e-2
    assert agg_rel(net.export_grads(), go) <= 1.5 * agg_rel(ge, go) + 2e-2

    initd = ON.init_params(ON.disc_param_shapes(), 3, 0.05)
    nz, mk = ON.make_disc_rand(rng, 1, S)
    gu = None

    def oracle_disc(emu):
        nonlocal gu
        ON.Emu.on = emu
        try:
            P = ON.to_torch(initd)
            xin = real_S.clone().requires_grad_(True)
            y = ON.disc_forward(P, xin, nz, mk)
            if gu is None:
                gu = torch.tensor(rng.standard_normal(y.shape), dtype=torch.float32)
            g = torch.autograd.grad(y, list(P.values()) + [xin], gu)
            return y.detach(), dict(zip(list(P.keys()) + ["dx"], g))
        finally:
            ON.Emu.on = False

    yo, go = oracle_disc(False)
    _, ge = oracle_disc(True)
    d = get_discriminator((S, S, S, 1), filters=64, use_dropout=True, use_input_noise=True, use_layer_noise=True, name='d')
    d.load(initd)
    tape = E.Tape()
    xv = E.Var(real_S.cuda())
    out = d.forward(tape, xv, training=True, noise=[t.cuda() for t in nz], masks=[m.cuda() for m in mk])
    d.zero_grad()
    tape.backward([(out, gu.cuda())], d.trainable_variables, wrt_vars=[xv])
    assert rel_l2(out.data.cpu(), yo) < 2e-2
    gk = d.export_grads()
    gk["dx"] = xv.grad.cpu().numpy()
    assert agg_rel(gk, go) <= 1.5 * agg_rel(ge, go) + 2e-2


def test_adam_update_and_second_step(cuda):
    """two consecutive optimizer steps.  Adam's first update is lr*sign(g) for every weight whatever the gradient
    magnitude, so near-zero gradient entries whose sign differs by rounding move weights differently: the step-2 losses
    are compared at 2e-1 (step 1 at 2e-2) and the weight displacement must correlate with the oracle's."""
    from oracle import losses as OL, nets as ON, step as OS
    from van_gan_b200.vangan import VanGan
    S, b = 32, 1
    real_I, real_S, init, _ = _setup(S, b, 1, 5, perturb=0.0)
    rng = np.random.default_rng(6)
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
            # step 2, measured against the oracle's 2.17 (gen_IS_loss): mma.sync path 2.22, tcgen05 paths 2.13-2.43 -- the fp32
            # summation order alone moves it by 12 %, so 2e-1 is what this ill-conditioned 32^3 case can pin
            tol = 2e-2 if it == 0 else 2e-1
            assert abs(res_k[k] - res_o[k]) <= tol * abs(res_o[k]) + 1e-3, (it, k, res_k[k], res_o[k])
    for name, net in gan.networks.items():
        w = net.export()
        num = den1 = den2 = 0.0
        for n in w:
            dk = torch.tensor(w[n] - init[name][n]).double().flatten()
            do = (P[name][n].detach() - torch.tensor(init[name][n])).double().flatten()
            num += float(dk @ do); den1 += float(dk @ dk); den2 += float(do @ do)
        assert num / (den1 * den2) ** 0.5 > 0.5, (name, num / (den1 * den2) ** 0.5)


def test_test_step_and_determinism(cuda):
    """test_step (training=False: no noise, no dropout) matches the oracle's inference-mode losses, and
    two train steps from identical state with explicit noise give identical losses."""
    from oracle import losses as OL, nets as ON, step as OS
    from van_gan_b200.vangan import VanGan
    S, b = 32, 1
    real_I, real_S, init, rand = _setup(S, b, 1, 9)
    res_o, _ = OS.compute_losses(OL.make_cfg(b, 1), {k: ON.to_torch(v, requires_grad=False) for k, v in init.items()},
                                 real_I, real_S, None)
    gan = VanGan(Args(S, b, 1), gen_i2s='resUnet', gen_s2i='resUnet')
    for k, net in gan.networks.items():
        net.load(init[k])
    res_t = gan.test_step(real_I, real_S)
    for k in OS.RESULT_KEYS:
        assert abs(res_t[k] - float(res_o[k])) <= 2e-2 * abs(float(res_o[k])) + 1e-4, k
    _, r1 = _cuda_step(S, b, 1, init, real_I, real_S, rand)
    _, r2 = _cuda_step(S, b, 1, init, real_I, real_S, rand)
    # the loss sums are accumulated with fp64 atomics: identical up to the summation order
    for k in r1:
        assert abs(r1[k] - r2[k]) <= 1e-12 * abs(r1[k]), k
