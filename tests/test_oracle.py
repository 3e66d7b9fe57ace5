"""CPU: pins the oracle.  The reference has no tests or golden vectors and TensorFlow is not installable
here, so the oracle is pinned by (1) an independent numpy/scipy restatement (bit-exact for the min/max
stencils), (2) analytic known answers, (3) fp64 finite-difference checks of its backward passes, (4) the
frozen fixtures under tests/golden/ (scripts/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import losses as OL, nets as ON, np_ref, step as OS

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("shape,iters", [((1, 24, 20, 28, 1), 6), ((2, 9, 33, 12, 1), 3), ((1, 16, 16, 16, 1), 15)])
def test_soft_skel_torch_numpy_dedup_bit_exact(shape, iters):
    rng = np.random.default_rng(0)
    x = torch.tensor(rng.random(shape), dtype=torch.float32)
    a, b = OL.soft_skel(x, iters), OL.soft_skel_dedup(x, iters)
    assert torch.equal(a, b)
    for n in range(shape[0]):
        assert np.array_equal(a[n, ..., 0].numpy(), np_ref.soft_skel(x[n, ..., 0].numpy(), iters))
    assert np.array_equal(OL.soft_erode(x)[0, ..., 0].numpy(), np_ref.soft_erode(x[0, ..., 0].numpy()))
    assert np.array_equal(OL.soft_dilate(x)[0, ..., 0].numpy(), np_ref.soft_dilate(x[0, ..., 0].numpy()))


def test_soft_skel_known_answers():
    # a constant volume has no skeleton; a one-voxel-wide straight line is its own skeleton
    c = torch.full((1, 12, 12, 12, 1), 0.7)
    assert float(OL.soft_skel(c, 5).abs().max()) == 0.0
    v = torch.zeros((1, 15, 15, 15, 1))
    v[0, 2:13, 7, 7, 0] = 1.0
    sk = OL.soft_skel(v, 5)
    assert torch.equal(sk, v)
    # a 3-voxel-thick bar erodes to its centre line
    v = torch.zeros((1, 17, 17, 17, 1))
    v[0, 2:15, 7:10, 7:10, 0] = 1.0
    sk = OL.soft_skel(v, 5)[0, ..., 0]
    assert float(sk[4:13, 8, 8].min()) == 1.0 and float(sk.sum()) <= 13.0 + 1e-6


def test_min_max_and_bce_and_ssim_against_numpy():
    rng = np.random.default_rng(1)
    t = torch.tensor(rng.random((2, 10, 11, 12, 1)), dtype=torch.float64)
    p = torch.tensor(rng.random((2, 10, 11, 12, 1)), dtype=torch.float64)
    n = OL.min_max_norm(t * 3 - 1)
    for i in range(2):
        assert np.allclose(n[i].numpy(), np_ref.min_max_norm((t[i] * 3 - 1).numpy()))
    s = 1 - OL.ssim_loss_3d(t, p)
    for i in range(2):
        assert np.abs(s[i, ..., 0].numpy() - np_ref.ssim_map(t[i, ..., 0].numpy(), p[i, ..., 0].numpy())).max() < 1e-12
    y, q = t.numpy(), np.clip(p.numpy(), 1e-7, 1 - 1e-7)
    ref = -(y * np.log(q + 1e-7) + (1 - y) * np.log(1 - q + 1e-7))
    assert np.allclose(OL.keras_bce(t, p).numpy(), ref[..., 0])
    # identical volumes: SSIM == 1 everywhere, Dice/clDice loss == 0 for a binary object with a skeleton
    assert float((1 - OL.ssim_loss_3d(t, t)).min()) > 1 - 1e-12


def test_reduce_mean_scalings_depend_on_replica_count():
    """loss_functions.py:7-22: axis=None -> mean over the LOCAL batch tensor / GLOBAL batch; per-sample axis -> sum of
    per-sample means / GLOBAL batch.  So BCE/SSIM terms change weight with the number of replicas at fixed global batch."""
    x = torch.arange(2 * 8, dtype=torch.float64).reshape(2, 2, 2, 2, 1)
    cfg = OL.make_cfg(4, 2)
    assert float(OL.reduce_mean(cfg, x)) == float(x.mean()) / 4
    assert float(OL.reduce_mean(cfg, x, axis=(1, 2, 3, 4))) == float(x.mean(dim=(1, 2, 3, 4)).sum()) / 4


def _fd_check(fn, x, eps=1e-6, n=6, seed=0):
    x = x.clone().requires_grad_(True)
    y = fn(x)
    (g,) = torch.autograd.grad(y, x)
    rng = np.random.default_rng(seed)
    for _ in range(n):
        v = torch.tensor(rng.standard_normal(x.shape), dtype=x.dtype)
        fd = (fn(x.detach() + eps * v) - fn(x.detach() - eps * v)) / (2 * eps)
        an = (g * v).sum()
        assert abs(float(fd) - float(an)) <= 1e-5 * max(1.0, abs(float(an))), (float(fd), float(an))


def test_backward_finite_differences_fp64():
    rng = np.random.default_rng(2)
    real = torch.tensor(rng.random((1, 8, 8, 8, 1)) * 2 - 1, dtype=torch.float64)
    cyc = torch.tensor(np.tanh(rng.standard_normal((1, 8, 8, 8, 1))), dtype=torch.float64)
    cfg = OL.make_cfg(2, 1)
    _fd_check(lambda c: OL.cycle_loss(cfg, real, c, typ="bce"), cyc)
    _fd_check(lambda c: OL.cycle_reconstruction(cfg, real, c), cyc)
    _fd_check(lambda c: OL.cycle_seg_loss(cfg, real, c, iters=3), cyc)
    # instance norm + conv through autograd (sanity of the layer restatement)
    x = torch.tensor(rng.standard_normal((1, 5, 5, 5, 8)), dtype=torch.float64)
    w = torch.tensor(rng.standard_normal((3, 3, 3, 8, 4)), dtype=torch.float64)
    g, b = torch.ones(8, dtype=torch.float64), torch.zeros(8, dtype=torch.float64)
    _fd_check(lambda t: (ON.conv3d(ON.reflect_pad(torch.relu(ON.instance_norm(t, g, b))), w, None, stride=2) ** 2).sum(), x)


def test_conv_padding_conventions():
    x = torch.arange(6 * 6 * 6, dtype=torch.float32).reshape(1, 6, 6, 6, 1)
    assert tuple(ON.reflect_pad(x).shape) == (1, 8, 8, 8, 1)
    assert float(ON.reflect_pad(x)[0, 0, 1, 1, 0]) == float(x[0, 1, 0, 0, 0])       # index -1 -> 1
    w1 = torch.ones((1, 1, 1, 1, 1))
    y = ON.conv3d(x, w1, None, stride=2, padding="same")                             # k1 s2: samples even indices
    assert torch.equal(y[0, :, :, :, 0], x[0, ::2, ::2, ::2, 0])
    w4 = torch.ones((4, 4, 4, 1, 1))
    y = ON.conv3d(torch.ones((1, 6, 6, 6, 1)), w4, None, stride=1, padding="same")   # TF 'same' k4: pad 1 before, 2 after
    assert float(y[0, 0, 0, 0, 0]) == 27.0 and float(y[0, 5, 5, 5, 0]) == 8.0 and float(y[0, 2, 2, 2, 0]) == 64.0


def test_param_counts_match_reference_models():
    assert sum(int(np.prod(s)) for s in ON.resunet_param_shapes().values()) == 9538929
    assert sum(int(np.prod(s)) for s in ON.disc_param_shapes().values()) == 11029953


def test_window_enumeration_matches_reference_counts():
    # 512x512x256, 128^3 windows, stride 64: 256 generator calls, 147 unique
    st = [(r, c, d) for r in np_ref.window_starts(512, 128, 64) for c in np_ref.window_starts(512, 128, 64)
          for d in np_ref.window_starts(256, 128, 64)]
    assert len(st) == 256 and len(set(st)) == 147
    # reference defaults complete=True, padFactor=0.25: padded 768x768x384 -> 864 calls, 605 unique
    st = [(r, c, d) for r in np_ref.window_starts(768, 128, 64) for c in np_ref.window_starts(768, 128, 64)
          for d in np_ref.window_starts(384, 128, 64)]
    assert len(st) == 864 and len(set(st)) == 605
    assert np_ref.window_starts(100, 64, 25) == [0, 25, 36]


def test_adam_matches_closed_form_first_step():
    w0 = torch.tensor([1.0, -2.0, 3.0])
    g = torch.tensor([0.5, -300.0, 0.0])
    P = {"a": w0.clone()}
    opt = OS.Adam(["a"])
    opt.apply(P, {"a": g})
    gc = g * (100.0 / max(float(g.norm()), 100.0))
    m, v = 0.5 * gc, 0.1 * gc * gc
    lr_t = 2e-4 * np.sqrt(1 - 0.9) / (1 - 0.5)
    assert torch.allclose(P["a"], w0 - lr_t * m / (v.sqrt() + 1e-7), rtol=1e-6)


def test_golden_fixtures_reproduced():
    torch.set_num_threads(1)
    z = np.load(os.path.join(GOLD, "losses_20.npz"))
    x01 = torch.tensor(z["x01"])
    assert np.array_equal(OL.soft_skel(x01, 5).numpy(), z["skel5"])
    assert np.array_equal(OL.soft_erode(x01).numpy(), z["erode"])
    cfg = OL.make_cfg(2, 2)
    r = torch.tensor(z["real"])
    fns = {"bce": lambda c: OL.cycle_loss(cfg, r, c, typ="bce"), "mse": lambda c: OL.cycle_loss(cfg, r, c, typ="mse"),
           "ssim": lambda c: OL.cycle_reconstruction(cfg, r, c), "seg": lambda c: OL.cycle_seg_loss(cfg, r, c, iters=5)}
    for k, fn in fns.items():
        c = torch.tensor(z["cycled"], requires_grad=True)
        l = fn(c)
        l.backward()
        assert abs(float(l) - float(z["val_" + k])) <= 1e-6 * abs(float(z["val_" + k]))
        assert np.allclose(c.grad.numpy(), z["grad_" + k], rtol=1e-4, atol=1e-9)
    z = np.load(os.path.join(GOLD, "stitch_40.npz"))
    gen = lambda a: np.tanh(1.5 * a - 0.3)
    a = np_ref.stitch_subvolumes(gen, z["vol"], (1, 16, 16, 16, 1), stride=(8, 8, 8), complete=True, padFactor=0.25)
    assert np.allclose(a, z["complete"], rtol=1e-6, atol=1e-4)
    b = np_ref.stitch_subvolumes(gen, z["vol"], (1, 16, 16, 16, 1), stride=(8, 8, 8), complete=False)
    assert np.array_equal(b, z["plain"])
    z = np.load(os.path.join(GOLD, "nets_32.npz"))
    Pd = ON.to_torch(ON.init_params(ON.disc_param_shapes(), 8, 0.05), requires_grad=False)
    yd = ON.disc_forward(Pd, torch.tensor(z["x"]))
    assert np.allclose(yd.numpy(), z["disc_out_inference"], rtol=1e-4, atol=1e-5)


def test_dp_emulation_sums_replica_gradients():
    """train_step_dp == manual shard-and-sum of replica_grads; result-dict entries are SUMMED (vangan.py:471-473)"""
    rng = np.random.default_rng(3)
    S, G = 32, 2
    shapes_g, shapes_d = ON.resunet_param_shapes(), ON.disc_param_shapes()
    P = {"gen_IS": ON.to_torch(ON.init_params(shapes_g, 1)), "gen_SI": ON.to_torch(ON.init_params(shapes_g, 2)),
         "disc_I": ON.to_torch(ON.init_params(shapes_d, 3)), "disc_S": ON.to_torch(ON.init_params(shapes_d, 4))}
    xI = torch.tensor(rng.random((G, S, S, S, 1)) * 2 - 1, dtype=torch.float32)
    xS = torch.tensor(rng.random((G, S, S, S, 1)) * 2 - 1, dtype=torch.float32)
    rands = [{k: ON.make_disc_rand(rng, 1, S) for k in ("S_real", "S_fake", "I_real", "I_fake")} for _ in range(G)]
    cfg = OL.make_cfg(G, G)
    res, grads = OS.train_step_dp(cfg, P, None, xI, xS, rands, iters=3)
    r0, g0, _ = OS.replica_grads(cfg, P, xI[:1], xS[:1], rands[0], iters=3)
    r1, g1, _ = OS.replica_grads(cfg, P, xI[1:], xS[1:], rands[1], iters=3)
    assert abs(res["total_IS_loss"] - float(r0["total_IS_loss"]) - float(r1["total_IS_loss"])) < 1e-5
    n = "dec0.cb1.conv.w"
    assert torch.allclose(grads["gen_IS"][n], g0["gen_IS"][n] + g1["gen_IS"][n], rtol=1e-5, atol=1e-7)


def test_vnet_oracle_matches_reference_parameter_count_and_shapes():
    """custom_vnet gen_IS variant (vangan.py:97-110): 25 888 737 trainable parameters, output shape =
    input shape, inference mode (masks=None) is deterministic and differs from a dropout draw."""
    from oracle import nets as ON
    assert sum(int(np.prod(s)) for s in ON.vnet_param_shapes(32, 4, 1).values()) == 25888737
    P = ON.to_torch(ON.init_params(ON.vnet_param_shapes(16, 3, 1), 3, 0.05), requires_grad=False)
    x = torch.tensor(np.random.default_rng(0).standard_normal((1, 16, 16, 16, 1)), dtype=torch.float32)
    y0 = ON.vnet_forward(P, x, 3)
    y1 = ON.vnet_forward(P, x, 3)
    ym = ON.vnet_forward(P, x, 3, masks=ON.make_vnet_masks(np.random.default_rng(1), 1, 16, 3))
    assert y0.shape == x.shape and torch.equal(y0, y1) and float(y0.abs().max()) <= 1.0
    assert not torch.allclose(y0, ym)


def test_conv_and_norm_against_independent_numpy_restatement():
    """The oracle's Conv3D / InstanceNormalization / UpSampling3D against a second restatement written with numpy only
    (sliding_window_view + einsum; explicit moments): Keras kernel layout (kd,kh,kw,Cin,Cout), VALID and strided, NDHWC."""
    from numpy.lib.stride_tricks import sliding_window_view
    rng = np.random.default_rng(21)
    for (cin, cout, k, s, sp) in [(3, 5, 3, 1, (6, 7, 8)), (4, 2, 4, 2, (10, 8, 12)), (2, 3, 1, 2, (5, 6, 7))]:
        x = rng.standard_normal((2,) + sp + (cin,))
        w = rng.standard_normal((k, k, k, cin, cout))
        b = rng.standard_normal(cout)
        win = sliding_window_view(x, (k, k, k), axis=(1, 2, 3))[:, ::s, ::s, ::s]     # (N, OD, OH, OW, Cin, kd, kh, kw)
        ref = np.einsum("ndhwcijk,ijkco->ndhwo", win, w) + b
        got = ON.conv3d(torch.tensor(x), torch.tensor(w), torch.tensor(b), stride=s).numpy()
        assert got.shape == ref.shape
        assert np.allclose(got, ref, rtol=1e-10, atol=1e-10), (cin, cout, k, s)
    x = rng.standard_normal((2, 4, 5, 6, 3)) * 2 + 1
    g, bt = rng.standard_normal(3), rng.standard_normal(3)
    mu = x.mean(axis=(1, 2, 3), keepdims=True)
    var = ((x - mu) ** 2).mean(axis=(1, 2, 3), keepdims=True)                          # biased, eps = 1e-3 (tfa default)
    ref = g * (x - mu) / np.sqrt(var + 1e-3) + bt
    got = ON.instance_norm(torch.tensor(x), torch.tensor(g), torch.tensor(bt)).numpy()
    assert np.allclose(got, ref, rtol=1e-10, atol=1e-10)
    up = ON.upsample2(torch.tensor(x)).numpy()
    assert np.array_equal(up, x.repeat(2, 1).repeat(2, 2).repeat(2, 3))


def test_instance_norm_gradient_properties_fp64():
    """Two analytic identities of y = gamma * xhat + beta, xhat = (x - mu) * rsqrt(var + eps), for ANY upstream gradient u:
    sum dx = 0 per (sample, channel), and  sum dx * xhat = gamma * rstd * (sum u * xhat) * eps / (var + eps)  -- the second one
    vanishes only for eps = 0, so it pins where the 1e-3 of tfa's InstanceNormalization enters."""
    rng = np.random.default_rng(22)
    x = torch.tensor(rng.standard_normal((2, 5, 4, 6, 3)), dtype=torch.float64, requires_grad=True)
    g = torch.tensor(1 + 0.3 * rng.standard_normal(3), dtype=torch.float64)
    b = torch.tensor(rng.standard_normal(3), dtype=torch.float64)
    up = torch.tensor(rng.standard_normal((2, 5, 4, 6, 3)), dtype=torch.float64)
    ON.instance_norm(x, g, b).backward(up)
    dx = x.grad
    assert float(dx.sum(dim=(1, 2, 3)).abs().max()) < 1e-12
    xd = x.detach()
    mu = xd.mean(dim=(1, 2, 3), keepdim=True)
    var = ((xd - mu) ** 2).mean(dim=(1, 2, 3), keepdim=True)
    rstd = 1.0 / torch.sqrt(var + 1e-3)
    xhat = (xd - mu) * rstd
    lhs = (dx * xhat).sum(dim=(1, 2, 3))
    rhs = (g * rstd * (up * xhat).sum(dim=(1, 2, 3), keepdim=True) * 1e-3 / (var + 1e-3)).reshape(lhs.shape)
    assert float((lhs - rhs).abs().max()) < 1e-12 * max(1.0, float(rhs.abs().max()))
    assert float(rhs.abs().max()) > 1e-6          # the identity is not vacuous


def test_keras_bce_and_lsgan_known_values():
    """Closed-form values of the Keras BCE branch (clip to [1e-7, 1-1e-7], log(p + 1e-7)) and of the LSGAN terms."""
    cfg = OL.make_cfg(1, 1)
    ones = torch.ones((1, 2, 2, 2, 1))
    zeros = torch.zeros((1, 2, 2, 2, 1))
    # generator: MSE(1, D(fake)) with D(fake) = 0 -> 1;  discriminator: 0.5 * (MSE(1, real=1) + MSE(0, fake=1)) = 0.5
    assert abs(float(OL.generator_loss_fn(cfg, zeros)) - 1.0) < 1e-7
    assert abs(float(OL.discriminator_loss_fn(cfg, ones, ones)) - 0.5) < 1e-7
    # bce cycle loss on a two-level volume: after per-sample min-max both tensors are {0,1}; identical -> -log(1 + 1e-7 - 1e-7)
    v = torch.tensor([0.0, 1.0] * 4).reshape(1, 2, 2, 2, 1) * 2 - 1
    same = float(OL.cycle_loss(cfg, v, v.clone(), typ="bce"))
    expect_same = -np.log(1 - 1e-7 + 1e-7) * 10.0
    assert abs(same - expect_same) < 1e-5
    flipped = float(OL.cycle_loss(cfg, v, -v, typ="bce"))                              # every voxel maximally wrong
    # fp32 like TF: y=1 voxels give -log(clip(0) + eps) = -log(2e-7); y=0 voxels -log(1 - clip(1) + eps), where 1 - 1e-7 rounds
    # to 1 - 2^-23 in fp32, i.e. -log(1.19e-7 + 1e-7): half the voxels each
    f = np.float32
    t1 = -np.log(f(1e-7) + f(1e-7))
    t0 = -np.log((f(1) - f(f(1) - f(1e-7))) + f(1e-7))
    expect_flip = 10.0 * 0.5 * (float(t1) + float(t0))
    assert abs(flipped - expect_flip) < 1e-4 * expect_flip


def test_adam_recursion_against_torch_adam_over_several_steps():
    """Independent second opinion on the optimizer recursion (moments, bias correction, step counter): with eps -> 0 Keras' form
    lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps) and torch.optim.Adam's lr/(1-b1^t) * m/(sqrt(v/(1-b2^t))+eps) are the same
    update, so five steps of the oracle's Adam must follow torch's on identical gradients (clipnorm out of reach)."""
    rng = np.random.default_rng(12)
    w0 = torch.tensor(rng.standard_normal(64), dtype=torch.float64)
    grads = [torch.tensor(rng.standard_normal(64), dtype=torch.float64) for _ in range(5)]
    P = {"a": w0.clone()}
    opt = OS.Adam(["a"], lr=2e-4, beta_1=0.5, beta_2=0.9, eps=1e-300, clipnorm=1e30)
    wt = w0.clone().requires_grad_(True)
    topt = torch.optim.Adam([wt], lr=2e-4, betas=(0.5, 0.9), eps=1e-300)
    for g in grads:
        opt.apply(P, {"a": g})
        wt.grad = g.clone()
        topt.step()
        assert torch.allclose(P["a"], wt.detach(), rtol=1e-12, atol=1e-15)
    # and the clip is per variable, on the aggregated gradient, before the moments see it (vangan.py:220-235: clipnorm=100)
    g = torch.tensor(rng.standard_normal(64) * 1e3, dtype=torch.float64)
    gc = OS.clip_by_norm(g, 100.0)
    assert abs(float(gc.norm()) - 100.0) < 1e-9 and torch.allclose(gc / gc.norm(), g / g.norm())
    assert torch.equal(OS.clip_by_norm(g * 1e-6, 100.0), g * 1e-6)


def test_batch_norm_against_torch_functional():
    """oracle.nets.batch_norm (Keras BatchNormalization, momentum 0.99, eps 1e-3) against torch.nn.functional.batch_norm: same
    normalised output and gradients in training mode, same moving averages (torch momentum = 1 - Keras momentum; both use the
    Bessel-corrected batch variance for the moving variance), same inference output."""
    import torch.nn.functional as F
    rng = np.random.default_rng(13)
    C = 6
    x = torch.tensor(rng.standard_normal((2, 3, 4, 5, C)) * 2 + 0.5, dtype=torch.float64, requires_grad=True)
    gamma = torch.tensor(rng.random(C) + 0.5, dtype=torch.float64, requires_grad=True)
    beta = torch.tensor(rng.standard_normal(C), dtype=torch.float64, requires_grad=True)
    gy = torch.tensor(rng.standard_normal((2, 3, 4, 5, C)), dtype=torch.float64)
    state = {"k.moving_mean": torch.zeros(C, dtype=torch.float64), "k.moving_variance": torch.ones(C, dtype=torch.float64)}
    y = ON.batch_norm(x, gamma, beta, state, "k", training=True)
    g_o = torch.autograd.grad((y * gy).sum(), [x, gamma, beta])
    rm, rv = torch.zeros(C, dtype=torch.float64), torch.ones(C, dtype=torch.float64)
    xt = x.detach().permute(0, 4, 1, 2, 3).clone().requires_grad_(True)
    gt, bt = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
    yt = F.batch_norm(xt, rm, rv, gt, bt, training=True, momentum=1.0 - ON.BN_MOMENTUM, eps=ON.BN_EPS)
    g_t = torch.autograd.grad((yt * gy.permute(0, 4, 1, 2, 3)).sum(), [xt, gt, bt])
    assert torch.allclose(y, yt.permute(0, 2, 3, 4, 1), rtol=1e-10, atol=1e-12)
    assert torch.allclose(g_o[0], g_t[0].permute(0, 2, 3, 4, 1), rtol=1e-9, atol=1e-12)
    assert torch.allclose(g_o[1], g_t[1], rtol=1e-9) and torch.allclose(g_o[2], g_t[2], rtol=1e-9)
    assert torch.allclose(state["k.moving_mean"], rm, rtol=1e-12) and torch.allclose(state["k.moving_variance"], rv, rtol=1e-12)
    yi = ON.batch_norm(x.detach(), gamma.detach(), beta.detach(), state, "k", training=False)
    yti = F.batch_norm(xt.detach(), rm, rv, gt.detach(), bt.detach(), training=False, eps=ON.BN_EPS)
    assert torch.allclose(yi, yti.permute(0, 2, 3, 4, 1), rtol=1e-10, atol=1e-12)


def test_instance_norm_and_transposed_conv_against_torch_functional():
    """tfa InstanceNormalization (eps 1e-3, biased variance) == torch instance_norm with the same eps; Conv3DTranspose k2 s2
    'same' == conv_transpose3d stride 2 with the Keras kernel (kd,kh,kw,Cout,Cin) permuted, checked by explicit index arithmetic."""
    import torch.nn.functional as F
    rng = np.random.default_rng(14)
    x = torch.tensor(rng.standard_normal((2, 4, 5, 6, 8)), dtype=torch.float64)
    gamma, beta = torch.tensor(rng.random(8) + 0.5, dtype=torch.float64), torch.tensor(rng.standard_normal(8), dtype=torch.float64)
    y = ON.instance_norm(x, gamma, beta)
    yt = F.instance_norm(x.permute(0, 4, 1, 2, 3), weight=gamma, bias=beta, eps=1e-3).permute(0, 2, 3, 4, 1)
    assert torch.allclose(y, yt, rtol=1e-10, atol=1e-12)
    xs = torch.tensor(rng.standard_normal((1, 2, 3, 2, 3)), dtype=torch.float64)
    w = torch.tensor(rng.standard_normal((2, 2, 2, 4, 3)), dtype=torch.float64)     # (kd, kh, kw, Cout, Cin)
    b = torch.tensor(rng.standard_normal(4), dtype=torch.float64)
    yc = ON.conv3d_transpose_k2s2(xs, w, b)
    ref = torch.zeros((1, 4, 6, 4, 4), dtype=torch.float64)
    for d in range(2):
        for h in range(3):
            for ww in range(2):
                for a in range(2):
                    for bb in range(2):
                        for c in range(2):
                            ref[0, 2 * d + a, 2 * h + bb, 2 * ww + c] = w[a, bb, c] @ xs[0, d, h, ww] + b
    assert torch.allclose(yc, ref, rtol=1e-12, atol=1e-12)
