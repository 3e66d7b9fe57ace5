"""GPU parity tests: every C-ABI kernel family against the CPU oracle on seeded inputs.

Tolerances: fp32 elementwise / stencil / loss kernels 1e-5 relative (soft_skel forward: bit-exact);
bf16 tensor-core convolutions: relative L2 <= 2e-2 against the fp32 oracle (in practice ~3e-3, the
bf16 rounding of the operands)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a = a.double().flatten().cpu()
    b = b.double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _bf(t):  # round through bf16 like the device path does for operands
    return t.to(torch.bfloat16).to(torch.float32)


# ----------------------------------------------------------------------------- soft skeleton
@pytest.mark.parametrize("shape,iters", [((1, 24, 20, 28), 6), ((2, 33, 17, 40), 3), ((1, 64, 64, 64), 15), ((1, 8, 8, 8), 0)])
def test_soft_skel_forward_bit_exact(cuda, shape, iters):
    from oracle import losses as OL
    from van_gan_b200 import clDice_func as K
    rng = np.random.default_rng(7)
    x = torch.tensor(rng.random(shape + (1,)), dtype=torch.float32)
    ref = OL.soft_skel(x, iters)
    out = K.soft_skel(x.cuda(), iters).cpu()
    assert torch.equal(out, ref)
    assert torch.equal(K.soft_erode(x.cuda()).cpu(), OL.soft_erode(x))


@pytest.mark.parametrize("shape,iters", [((1, 20, 24, 36), 5), ((2, 16, 16, 16), 2), ((1, 40, 40, 40), 15)])
def test_soft_skel_backward(cuda, shape, iters):
    from oracle import losses as OL
    from van_gan_b200 import clDice_func as K
    rng = np.random.default_rng(8)
    x = torch.tensor(rng.random(shape + (1,)), dtype=torch.float32, requires_grad=True)
    g = torch.tensor(rng.standard_normal(shape + (1,)), dtype=torch.float32)
    OL.soft_skel(x, iters).backward(g)
    skel, bwd = K.soft_skel_with_grad(x.detach().cuda(), iters)
    dx = bwd(g.cuda()).cpu()
    assert rel_l2(dx, x.grad) < 1e-5
    assert float((dx - x.grad).abs().max()) < 1e-5 * float(x.grad.abs().max()) + 1e-6


@pytest.mark.parametrize("scale", [1e-20, 1e-6, 1e20])
def test_soft_skel_backward_gradient_scale(cuda, scale):
    """The routing kernel accumulates in fixed point with a scale taken from the largest value of each level: the result must be
    as accurate for tiny and huge incoming gradients as for O(1) ones, and bit-identical from call to call."""
    from oracle import losses as OL
    from van_gan_b200 import clDice_func as K
    shape, iters = (2, 18, 26, 34), 7
    rng = np.random.default_rng(28)
    x = torch.tensor(rng.random(shape + (1,)), dtype=torch.float32, requires_grad=True)
    g = torch.tensor(rng.standard_normal(shape + (1,)), dtype=torch.float32)
    g[0, 3, 4, 5, 0] = 300.0                                        # one dominant element sets the scale for everything else
    OL.soft_skel(x, iters).backward(g)
    _skel, bwd = K.soft_skel_with_grad(x.detach().cuda(), iters)
    dx = bwd((g * scale).cuda())
    assert torch.equal(dx, bwd((g * scale).cuda()))
    dx = (dx.double() / scale).float().cpu()
    assert rel_l2(dx, x.grad) < 1e-5
    assert float((dx - x.grad).abs().max()) < 1e-5 * float(x.grad.abs().max())


@pytest.mark.parametrize("shape,iters,levels", [((2, 37, 21, 45), 4, 4), ((1, 70, 18, 30), 6, 2), ((3, 9, 40, 33), 3, 8)])
def test_soft_skel_backward_with_ties(cuda, shape, iters, levels, monkeypatch):
    """Quantised volumes (segmentation-like plateaus): almost every min / max window holds several equal extrema, so the gradient
    routing is decided by the tie rule in every level.  The policy (oracle/__init__.py: one winner, the first extremum in the window's
    scan order) is implemented twice -- the z-marching routing kernel (separable first-extremum search, shared-memory scatter) and the
    tile kernel (VG_SKEL_BWD=tile: explicit window scans, gather form) -- and the two must agree; the forward has no tie ambiguity and
    is compared with the oracle.  Ragged tile / z-chunk extents."""
    from oracle import losses as OL
    from van_gan_b200 import clDice_func as K
    rng = np.random.default_rng(18)
    xq = np.round(rng.random(shape + (1,)) * levels) / levels
    x = torch.tensor(xq, dtype=torch.float32)
    g = torch.tensor(rng.standard_normal(shape + (1,)), dtype=torch.float32).cuda()
    skel, bwd = K.soft_skel_with_grad(x.cuda(), iters)
    assert torch.equal(skel.cpu(), OL.soft_skel(x, iters))
    monkeypatch.delenv("VG_SKEL_BWD", raising=False)
    dx_march = bwd(g).clone()
    monkeypatch.setenv("VG_SKEL_BWD", "tile")
    dx_tile = bwd(g).clone()
    assert float(dx_tile.abs().max()) > 0
    assert rel_l2(dx_march.cpu(), dx_tile.cpu()) < 1e-6
    # conservation: every level routes each incoming gradient to exactly one voxel, so no gradient mass is created or lost by ties
    assert float((dx_march - dx_tile).abs().max()) < 1e-5 * float(dx_tile.abs().max())


# ----------------------------------------------------------------------------- losses
class _Cfg:
    def __init__(self, G=2, nd=1):
        self.global_batch_size, self.n_devices = G, nd
        self.lambda_cycle, self.lambda_reconstruction, self.lambda_topology = 10.0, 5.0, 5.0
        self.loss_ctx = None


@pytest.mark.parametrize("S,N", [(24, 2), (32, 1)])
def test_cycle_losses_forward_backward(cuda, S, N):
    from oracle import losses as OL
    from van_gan_b200 import engine as E, loss_functions as LF
    rng = np.random.default_rng(11)
    real = torch.tensor(rng.random((N, S, S, S, 1)) * 2 - 1, dtype=torch.float32)
    cyc0 = torch.tensor(np.tanh(rng.standard_normal((N, S, S, S, 1))), dtype=torch.float32)
    cfg = _Cfg(G=2 * N, nd=2)
    ocfg = OL.make_cfg(2 * N, 2)
    cases = {
        "bce": (lambda c: OL.cycle_loss(ocfg, real, c, typ="bce"), lambda r, c: LF.cycle_loss(cfg, r, c, typ="bce")),
        "mse": (lambda c: OL.cycle_loss(ocfg, real, c, typ="mse"), lambda r, c: LF.cycle_loss(cfg, r, c, typ="mse")),
        "ssim": (lambda c: OL.cycle_reconstruction(ocfg, real, c), lambda r, c: LF.cycle_reconstruction(cfg, r, c)),
        "seg": (lambda c: OL.cycle_seg_loss(ocfg, real, c, iters=5), lambda r, c: LF.cycle_seg_loss(cfg, r, c, iters=5)),
    }
    for name, (ofn, kfn) in cases.items():
        c = cyc0.clone().requires_grad_(True)
        lo = ofn(c)
        lo.backward()
        cfg.loss_ctx = LF.LossContext()
        rv, cv = E.Var(real.cuda()), E.Var(cyc0.cuda())
        s = kfn(rv, cv)
        assert abs(float(s) - float(lo)) <= 1e-5 * abs(float(lo)) + 1e-7, name
        seeds = s.seeds()
        g = sum(gg for v, gg in seeds if v is cv).cpu()
        assert rel_l2(g, c.grad) < 2e-5, (name, rel_l2(g, c.grad))


def test_lsgan_losses(cuda):
    from oracle import losses as OL
    from van_gan_b200 import engine as E, loss_functions as LF
    rng = np.random.default_rng(12)
    a = torch.tensor(rng.standard_normal((2, 4, 4, 4, 1)), dtype=torch.float32)
    b = torch.tensor(rng.standard_normal((2, 4, 4, 4, 1)), dtype=torch.float32)
    cfg, ocfg = _Cfg(G=4), OL.make_cfg(4)
    cfg.loss_ctx = LF.LossContext()
    ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    lo = OL.discriminator_loss_fn(ocfg, ar, br) + OL.generator_loss_fn(ocfg, br)
    lo.backward()
    av, bv = E.Var(a.cuda()), E.Var(b.cuda())
    s = LF.discriminator_loss_fn(cfg, av, bv) + LF.generator_loss_fn(cfg, bv)
    assert abs(float(s) - float(lo)) < 1e-6 * abs(float(lo))
    seeds = s.seeds()
    ga = sum(g for v, g in seeds if v is av).cpu()
    gb = sum(g for v, g in seeds if v is bv).cpu()
    assert rel_l2(ga, ar.grad) < 1e-6 and rel_l2(gb, br.grad) < 1e-6


# ----------------------------------------------------------------------------- instance norm
@pytest.mark.parametrize("C,S,act,pad,dt", [(16, 12, 1, (1, 1, 1), torch.float32), (48, 10, 0, (0, 0, 0), torch.float32),
                                            (64, 8, 2, (1, 2, 0), torch.float32), (32, 9, 2, (1, 1, 1), torch.bfloat16),
                                            (512, 4, 2, (1, 1, 0), torch.float32)])
def test_instnorm_forward_backward(cuda, C, S, act, pad, dt):
    import torch.nn.functional as F
    from oracle import nets as ON
    from van_gan_b200 import engine as E
    from van_gan_b200.resunet_model import resunet_param_shapes  # noqa: F401
    rng = np.random.default_rng(13)
    N = 2
    x = torch.tensor(rng.standard_normal((N, S, S + 1, S + 2, C)) * 1.5 + 0.7, dtype=torch.float32)
    res = torch.tensor(rng.standard_normal(x.shape), dtype=torch.float32)
    gamma = torch.tensor(1 + 0.2 * rng.standard_normal(C), dtype=torch.float32)
    beta = torch.tensor(0.2 * rng.standard_normal(C), dtype=torch.float32)
    drop = torch.tensor((rng.random((N, 1, 1, 1, C)) > 0.2) / 0.8, dtype=torch.float32)
    if dt == torch.bfloat16:
        x, res = _bf(x), _bf(res)
    pshape = (N, S + pad[0] + pad[1], S + 1 + pad[0] + pad[1], S + 2 + pad[0] + pad[1], C)
    noise_shape = pshape if pad[2] == 1 else x.shape
    noise = torch.tensor(0.1 * rng.standard_normal(noise_shape), dtype=torch.float32)
    gout = torch.tensor(rng.standard_normal(pshape), dtype=torch.float32)
    if dt == torch.bfloat16:
        gout = _bf(gout)

    xr, rr = x.clone().requires_grad_(True), res.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = ON.instance_norm(xr, gr, br)
    y = torch.relu(y) if act == 1 else (F.leaky_relu(y, 0.2) if act == 2 else y)
    y = y * drop + rr
    if pad[2] == 1 and pad[0]:
        y = ON.reflect_pad(y) + noise
    else:
        y = y + noise
        if pad[0] or pad[1]:
            y = F.pad(y.permute(0, 4, 1, 2, 3), (pad[0], pad[1]) * 3).permute(0, 2, 3, 4, 1)
    y.backward(gout)

    class Net:
        pass
    from collections import OrderedDict
    net = E.Network("t", OrderedDict([("n.gamma", (C,)), ("n.beta", (C,))]))
    net.load({"n.gamma": gamma.numpy(), "n.beta": beta.numpy()})
    layer = E.InstanceNorm(net, "n", C)
    tape = E.Tape()
    xv, rv = E.Var(x.to(dt).cuda()), E.Var(res.to(dt).cuda())
    out = layer(tape, xv, act=act, residual=rv, pad=pad, drop=drop.reshape(-1).cuda(), noise=noise.cuda())
    tol = 1e-5 if dt == torch.float32 else 1.5e-2
    assert rel_l2(out.data.float(), y.detach()) < tol
    tape.backward([(out, gout.to(dt).cuda())], net.trainable_variables, wrt_vars=[xv, rv])
    tolb = 2e-5 if dt == torch.float32 else 2e-2
    assert rel_l2(xv.grad.float(), xr.grad) < tolb
    assert rel_l2(rv.grad.float(), rr.grad) < tolb
    assert rel_l2(net.params["n.gamma"].grad, gr.grad) < tolb
    assert rel_l2(net.params["n.beta"].grad, br.grad) < tolb


# ----------------------------------------------------------------------------- convolutions
CONV_CASES = [
    # (Cin, Cout, K, stride, spatial(in, padded), N)
    (16, 16, 3, 1, (10, 11, 14), 2), (48, 16, 3, 1, (9, 10, 18), 1), (16, 32, 3, 2, (13, 11, 19), 2),
    (32, 32, 3, 1, (6, 6, 10), 1), (96, 32, 3, 1, (8, 8, 10), 1), (64, 128, 3, 2, (9, 9, 9), 1),
    (256, 256, 3, 1, (4, 4, 4), 2), (384, 128, 3, 1, (6, 6, 6), 1), (192, 64, 3, 1, (7, 6, 10), 1),
    (16, 32, 1, 2, (12, 12, 16), 1), (48, 16, 1, 1, (8, 8, 8), 2), (128, 256, 1, 2, (4, 4, 4), 1),
    (64, 128, 4, 2, (10, 10, 18), 1), (128, 256, 4, 2, (10, 10, 10), 1), (256, 512, 4, 1, (7, 7, 7), 1),
    (1, 16, 3, 1, (10, 10, 12), 2), (1, 64, 4, 2, (18, 18, 18), 1), (1, 16, 1, 1, (8, 8, 8), 1),
    (512, 1, 3, 1, (6, 6, 6), 2), (16, 1, 1, 1, (8, 8, 9), 2),
    # streaming kernels of conv_small.cu: ragged voxel counts, every template instance, odd extents for the fused-parity dgrad
    (48, 16, 1, 1, (5, 7, 9), 1), (96, 32, 1, 1, (8, 8, 10), 1), (32, 64, 1, 2, (8, 9, 12), 1), (16, 16, 1, 1, (7, 5, 3), 2),
    (32, 32, 1, 1, (6, 6, 7), 1), (1, 64, 4, 2, (19, 21, 40), 2), (512, 1, 3, 1, (5, 6, 9), 1), (16, 1, 1, 1, (33, 5, 7), 1),
    # more work items than SMs: persistent tcgen05 CTAs re-use their (zeroed) TMEM buffers and run several d-march bricks
    (16, 16, 3, 1, (34, 34, 66), 3), (32, 32, 3, 1, (34, 34, 34), 5),
    # fused parity-class stride-2 dgrad (Cin 16 / 32): two N blocks, k4 taps, odd extents
    (32, 64, 3, 2, (11, 9, 13), 1), (16, 32, 4, 2, (10, 12, 14), 1), (16, 32, 3, 2, (34, 18, 35), 2), (128, 256, 3, 2, (9, 7, 11), 2),
    # 7^3 convolutions of the 'resnet' generator (generator.py:38,67): single-channel input / single-channel output (csrc/conv_k7.cu)
    (1, 32, 7, 1, (14, 13, 16), 2), (32, 1, 7, 1, (13, 14, 15), 1), (16, 1, 7, 1, (9, 9, 12), 2),
    # k4 s1 after UpSampling3D (building_blocks.upsample) at the resnet generator's widths
    (256, 128, 4, 1, (7, 7, 7), 1), (128, 64, 4, 1, (11, 11, 11), 1), (64, 32, 4, 1, (19, 11, 13), 2), (32, 64, 3, 2, (15, 15, 17), 1),
]


# every convolution shape of the 128^3 train step at its FULL spatial size (BASELINE config 2), against the oracle's conv3d:
# persistent CTAs run tens of work items each, the d-march / d-split / fused-class paths see their real grids
FULL_CASES = [
    (16, 16, 3, 1, (130, 130, 130), 1), (48, 16, 3, 1, (130, 130, 130), 1), (16, 32, 3, 2, (130, 130, 130), 1),
    (32, 32, 3, 1, (66, 66, 66), 1), (96, 32, 3, 1, (66, 66, 66), 1), (32, 64, 3, 2, (66, 66, 66), 1),
    (64, 64, 3, 1, (34, 34, 34), 2), (192, 64, 3, 1, (34, 34, 34), 1), (64, 128, 3, 2, (34, 34, 34), 2),
    (128, 128, 3, 1, (18, 18, 18), 2), (384, 128, 3, 1, (18, 18, 18), 1), (128, 256, 3, 2, (18, 18, 18), 2),
    (256, 256, 3, 1, (10, 10, 10), 8), (64, 128, 4, 2, (66, 66, 66), 1), (128, 256, 4, 2, (34, 34, 34), 2),
    (256, 512, 4, 1, (19, 19, 19), 2), (1, 16, 3, 1, (130, 130, 130), 1), (1, 64, 4, 2, (130, 130, 130), 1),
    (48, 16, 1, 1, (128, 128, 128), 1), (16, 32, 1, 2, (128, 128, 128), 1), (16, 1, 1, 1, (128, 128, 128), 1),
    (512, 1, 3, 1, (18, 18, 18), 4), (16, 16, 3, 1, (130, 130, 130), 4),
]


@pytest.mark.parametrize("Cin,Cout,K,stride,sp,N", FULL_CASES)
def test_conv3d_full_size_vs_oracle(cuda, Cin, Cout, K, stride, sp, N):
    # oracle in fp64 (exact reference for sums over 2 M voxels); the kernels accumulate in fp32: 2e-4 instead of 1e-5 on dw / dbias
    _conv_case(Cin, Cout, K, stride, sp, N, odt=torch.float64, wtol=2e-4)


@pytest.mark.parametrize("Cin,Cout,K,stride,sp,N", CONV_CASES)
def test_conv3d_fwd_dgrad_wgrad(cuda, Cin, Cout, K, stride, sp, N):
    _conv_case(Cin, Cout, K, stride, sp, N)


def _conv_case(Cin, Cout, K, stride, sp, N, odt=torch.float32, wtol=1e-5):
    from collections import OrderedDict
    from oracle import nets as ON
    from van_gan_b200 import engine as E
    from van_gan_b200._lib import ACT_NONE, ACT_TANH
    rng = np.random.default_rng(Cin * 1000 + Cout + K)
    act = ACT_TANH if (Cout == 1 and K == 1) else ACT_NONE
    x = torch.tensor(rng.standard_normal((N,) + sp + (Cin,)), dtype=torch.float32)
    w = torch.tensor(ON.he_normal(rng, (K, K, K, Cin, Cout)), dtype=torch.float32)
    b = torch.tensor(0.1 * rng.standard_normal(Cout), dtype=torch.float32)
    xd = x if Cin == 1 else _bf(x)
    wd = w if Cin == 1 else _bf(w)      # tensor-core layers use the bf16 operand copy of the weights
    xr, wr, br = [t.to(odt).clone().requires_grad_(True) for t in (xd, wd, b)]
    y = ON.conv3d(xr, wr, br, stride=stride)
    if act == ACT_TANH:
        y = torch.tanh(y)
    gy = torch.tensor(rng.standard_normal(y.shape), dtype=torch.float32)
    if Cout != 1:
        gy = _bf(gy)
    y.backward(gy.to(odt))

    net = E.Network("t", OrderedDict([("c.w", tuple(w.shape)), ("c.b", (Cout,))]))
    net.load({"c.w": wd.numpy(), "c.b": b.numpy()})
    layer = E.Conv3D(net, "c", K, stride, Cin, Cout, act=act)
    net.repack()
    tape = E.Tape()
    xv = E.Var(x.cuda() if Cin == 1 else x.to(torch.bfloat16).cuda())
    out = layer(tape, xv)
    assert out.shape == tuple(y.shape)
    # operands are bf16-exact, so the only differences are fp32 summation order and the final bf16 store:
    # compare against the oracle rounded the same way (a wrong tap / border / stride shows up as >> 1e-3)
    yref = y.detach().float() if Cout == 1 else _bf(y.detach().float())
    assert rel_l2(out.data.float(), yref) < 1e-3, "fwd"
    assert rel_l2(out.data.float(), y.detach()) < 2e-2          # north_star tolerance vs the fp32 oracle
    tape.backward([(out, gy.cuda() if Cout == 1 else gy.to(torch.bfloat16).cuda())], net.trainable_variables, wrt_vars=[xv])
    dxref = xr.grad.float() if Cin == 1 else _bf(xr.grad.float())
    assert rel_l2(xv.grad.float(), dxref) < (5e-3 if Cin == 1 else 1e-3), "dgrad"   # Cin==1: dgrad runs on bf16 weights
    assert rel_l2(net.params["c.w"].grad, wr.grad) < wtol, "wgrad"
    assert rel_l2(net.params["c.b"].grad, br.grad) < wtol, "bias grad"


def test_upsample_concat_and_pad(cuda):
    from oracle import nets as ON
    from van_gan_b200 import engine as E
    rng = np.random.default_rng(3)
    lo = _bf(torch.tensor(rng.standard_normal((2, 3, 4, 5, 32)), dtype=torch.float32))
    sk = _bf(torch.tensor(rng.standard_normal((2, 6, 8, 10, 16)), dtype=torch.float32))
    lr, sr = lo.clone().requires_grad_(True), sk.clone().requires_grad_(True)
    y = torch.cat([ON.upsample2(lr), sr], dim=-1)
    g = _bf(torch.tensor(rng.standard_normal(y.shape), dtype=torch.float32))
    y.backward(g)
    tape = E.Tape()
    lv, sv = E.Var(lo.to(torch.bfloat16).cuda()), E.Var(sk.to(torch.bfloat16).cuda())
    out = E.upsample_concat(tape, lv, sv)
    assert torch.equal(out.data.float().cpu(), y.detach())
    tape.backward([(out, g.to(torch.bfloat16).cuda())], [], wrt_vars=[lv, sv])
    assert rel_l2(lv.grad.float(), lr.grad) < 1e-2
    assert torch.equal(sv.grad.float().cpu(), sr.grad)

    x = torch.tensor(rng.standard_normal((2, 5, 6, 7, 1)), dtype=torch.float32)
    nz = torch.tensor(rng.standard_normal((2, 7, 8, 9, 1)), dtype=torch.float32)
    xr = x.clone().requires_grad_(True)
    yp = ON.reflect_pad(xr) + nz
    gp = torch.tensor(rng.standard_normal(yp.shape), dtype=torch.float32)
    yp.backward(gp)
    tape = E.Tape()
    xv = E.Var(x.cuda())
    o = E.pad_noise(tape, xv, noise=nz.cuda())
    assert torch.allclose(o.data.cpu(), yp.detach(), atol=1e-6)
    tape.backward([(o, gp.cuda())], [], wrt_vars=[xv])
    assert torch.allclose(xv.grad.cpu(), xr.grad, atol=1e-5)


def test_clip_adam(cuda):
    from collections import OrderedDict
    from oracle import step as OS
    from van_gan_b200 import engine as E
    rng = np.random.default_rng(5)
    shapes = OrderedDict([("a.w", (3, 3, 3, 4, 8)), ("a.b", (8,)), ("b.gamma", (16,)), ("c.w", (1, 1, 1, 64, 64))])
    net = E.Network("t", shapes)
    init = {k: rng.standard_normal(s).astype(np.float32) for k, s in shapes.items()}
    net.load(init)
    P = OrderedDict((k, torch.tensor(v)) for k, v in init.items())
    opt = OS.Adam(list(shapes))
    for it in range(3):
        g = {k: (rng.standard_normal(s) * (40.0 if k == "c.w" else 1.0)).astype(np.float32) for k, s in shapes.items()}
        for k in shapes:
            net.params[k].grad.copy_(torch.tensor(g[k]).cuda())
        net.adam_step()
        opt.apply(P, {k: torch.tensor(v) for k, v in g.items()})
        for k in shapes:
            assert torch.allclose(net.params[k].w.cpu(), P[k], rtol=1e-5, atol=1e-6), (it, k)


@pytest.mark.parametrize("sp,C,S", [(1, 16, 12), (1, 48, 8), (2, 16, 12), (2, 64, 6)])
def test_instnorm_specialised_generator_cases(cuda, sp, C, S):
    """The two compile-time-specialised InstanceNorm instances of the generator hot path (bf16, no dropout / noise):
    sp=1 InstanceNorm -> ReLU -> ReflectionPadding3D (resunet_model.py:42-66); sp=2 InstanceNorm + residual Add
    (resunet_model.py:96-100,133-143).  Forward and all gradients against the fp32 oracle on bf16-valued inputs."""
    from collections import OrderedDict
    from oracle import nets as ON
    from van_gan_b200 import engine as E
    from van_gan_b200._lib import ACT_NONE, ACT_RELU, PAD_REFLECT, PAD_ZERO
    rng = np.random.default_rng(17 + sp)
    N = 2
    x = _bf(torch.tensor(rng.standard_normal((N, S, S + 1, S + 2, C)) * 1.5 + 0.7, dtype=torch.float32))
    res = _bf(torch.tensor(rng.standard_normal(x.shape), dtype=torch.float32))
    gamma = torch.tensor(1 + 0.2 * rng.standard_normal(C), dtype=torch.float32)
    beta = torch.tensor(0.2 * rng.standard_normal(C), dtype=torch.float32)
    xr, rr = x.clone().requires_grad_(True), res.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = ON.instance_norm(xr, gr, br)
    y = ON.reflect_pad(torch.relu(y)) if sp == 1 else y + rr
    gout = _bf(torch.tensor(rng.standard_normal(y.shape), dtype=torch.float32))
    y.backward(gout)
    net = E.Network("t", OrderedDict([("n.gamma", (C,)), ("n.beta", (C,))]))
    net.load({"n.gamma": gamma.numpy(), "n.beta": beta.numpy()})
    layer = E.InstanceNorm(net, "n", C)
    tape = E.Tape()
    xv, rv = E.Var(x.to(torch.bfloat16).cuda()), E.Var(res.to(torch.bfloat16).cuda())
    if sp == 1:
        out = layer(tape, xv, act=ACT_RELU, pad=(1, 1, PAD_REFLECT))
    else:
        out = layer(tape, xv, act=ACT_NONE, residual=rv)
    assert rel_l2(out.data.float(), y.detach()) < 1.5e-2
    tape.backward([(out, gout.to(torch.bfloat16).cuda())], net.trainable_variables, wrt_vars=[xv, rv] if sp == 2 else [xv])
    assert rel_l2(xv.grad.float(), xr.grad) < 2e-2
    if sp == 2:
        assert rel_l2(rv.grad.float(), rr.grad) < 2e-2
    assert rel_l2(net.params["n.gamma"].grad, gr.grad) < 2e-2
    assert rel_l2(net.params["n.beta"].grad, br.grad) < 2e-2
