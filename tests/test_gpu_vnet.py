"""GPU parity of the V-Net generator variant (vnet_model.py:80-146,149-268 with the gen_IS arguments at
vangan.py:97-110) and of its layout kernels against the CPU oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float(torch.linalg.vector_norm(a - b) / (torch.linalg.vector_norm(b) + 1e-30))


def _bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("up,pad,mode,C0,C1", [(2, 1, 0, 16, 0), (1, 1, 1, 16, 8), (2, 0, 0, 8, 16), (1, 1, 1, 0, 16)])
def test_gather_pad_fwd_bwd(cuda, up, pad, mode, C0, C1):
    """UpSampling3D / concatenate / ReflectionPadding3D / zero 'same' padding in one pass, and its adjoint."""
    import torch.nn.functional as F
    from van_gan_b200 import engine as E
    rng = np.random.default_rng(7)
    N, D, H, W = 2, 6, 4, 8
    a = _bf(torch.tensor(rng.standard_normal((N, D // up, H // up, W // up, C0)), dtype=torch.float32)) if C0 else None
    b = _bf(torch.tensor(rng.standard_normal((N, D, H, W, C1)), dtype=torch.float32)) if C1 else None
    ar = a.clone().requires_grad_(True) if C0 else None
    br = b.clone().requires_grad_(True) if C1 else None
    parts = []
    if C0:
        parts.append(ar.repeat_interleave(up, 1).repeat_interleave(up, 2).repeat_interleave(up, 3))
    if C1:
        parts.append(br)
    y = torch.cat(parts, dim=-1)
    if pad:
        y = F.pad(y.permute(0, 4, 1, 2, 3), (1,) * 6, mode="reflect" if mode == 1 else "constant").permute(0, 2, 3, 4, 1)
    g = _bf(torch.tensor(rng.standard_normal(y.shape), dtype=torch.float32))
    y.backward(g)
    tape = E.Tape()
    av = E.Var(a.to(torch.bfloat16).cuda()) if C0 else None
    bv = E.Var(b.to(torch.bfloat16).cuda()) if C1 else None
    out = E.gather_pad(tape, av, bv, up=up, pad=pad, mode=mode)
    assert torch.equal(out.data.float().cpu(), y.detach())            # pure data movement: exact
    tape.backward([(out, g.to(torch.bfloat16).cuda())], [], wrt_vars=[v for v in (av, bv) if v is not None])
    if C0:
        assert rel_l2(av.grad.float(), ar.grad) < 4e-3                # bf16 store of an up^3 * fold sum
    if C1:
        assert rel_l2(bv.grad.float(), br.grad) < 4e-3


@pytest.mark.parametrize("pad,mode", [(1, 1), (0, 0)])
def test_maxpool_pad_fwd_bwd(cuda, pad, mode):
    import torch.nn.functional as F
    from van_gan_b200 import engine as E
    rng = np.random.default_rng(9)
    N, D, H, W, C = 2, 8, 4, 6, 16
    x = _bf(torch.tensor(rng.standard_normal((N, D, H, W, C)), dtype=torch.float32))
    xr = x.clone().requires_grad_(True)
    y = F.max_pool3d(xr.permute(0, 4, 1, 2, 3), 2)
    if pad:
        y = F.pad(y, (1,) * 6, mode="reflect")
    y = y.permute(0, 2, 3, 4, 1)
    g = _bf(torch.tensor(rng.standard_normal(y.shape), dtype=torch.float32))
    y.backward(g)
    tape = E.Tape()
    xv = E.Var(x.to(torch.bfloat16).cuda())
    out = E.maxpool_pad(tape, xv, pad=pad, mode=mode)
    assert torch.equal(out.data.float().cpu(), y.detach())
    tape.backward([(out, g.to(torch.bfloat16).cuda())], [], wrt_vars=[xv])
    assert rel_l2(xv.grad.float(), xr.grad) < 4e-3


@pytest.mark.parametrize("C,S,pad,dt,tol", [(16, 10, (1, 1, 1), torch.float32, 2e-5), (32, 8, (0, 0, 0), torch.float32, 2e-5),
                                             (16, 10, (1, 1, 1), torch.bfloat16, 2e-2)])
def test_instnorm_on_relu_input(cuda, C, S, pad, dt, tol):
    """Conv3D(activation='relu') -> InstanceNormalization -> SpatialDropout3D -> ReflectionPadding3D (vnet_model.py:116-132):
    the ReLU is applied on load by the norm kernels and its mask on dx.  fp32 storage: 2e-5."""
    from collections import OrderedDict
    from oracle import nets as ON
    from van_gan_b200 import engine as E
    rng = np.random.default_rng(21)
    N = 2
    x = torch.tensor(rng.standard_normal((N, S, S + 1, S + 2, C)) * 1.3 + 0.2, dtype=torch.float32)
    gamma = torch.tensor(1 + 0.2 * rng.standard_normal(C), dtype=torch.float32)
    beta = torch.tensor(0.2 * rng.standard_normal(C), dtype=torch.float32)
    drop = torch.tensor((rng.random((N, 1, 1, 1, C)) > 0.5) / 0.5, dtype=torch.float32)
    if dt == torch.bfloat16:
        x = _bf(x)
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = ON.instance_norm(torch.relu(xr), gr, br) * drop
    if pad[0]:
        y = ON.reflect_pad(y)
    gout = torch.tensor(rng.standard_normal(y.shape), dtype=torch.float32)
    if dt == torch.bfloat16:
        gout = _bf(gout)
    y.backward(gout)
    net = E.Network("t", OrderedDict([("n.gamma", (C,)), ("n.beta", (C,))]))
    net.load({"n.gamma": gamma.numpy(), "n.beta": beta.numpy()})
    layer = E.InstanceNorm(net, "n", C)
    tape = E.Tape()
    xv = E.Var(x.to(dt).cuda())
    out = layer(tape, xv, drop=drop.reshape(-1).cuda(), pad=pad, relu_input=True)
    assert rel_l2(out.data.float(), y.detach()) < (1e-5 if dt == torch.float32 else 1.5e-2)
    tape.backward([(out, gout.to(dt).cuda())], net.trainable_variables, wrt_vars=[xv])
    assert rel_l2(xv.grad.float(), xr.grad) < tol
    assert rel_l2(net.params["n.gamma"].grad, gr.grad) < tol
    assert rel_l2(net.params["n.beta"].grad, br.grad) < tol


def _vnet_case(S, filters, L, N, seed=1):
    from oracle import nets as ON
    rng = np.random.default_rng(seed)
    shapes = ON.vnet_param_shapes(filters, L, 1)
    init = ON.init_params(shapes, 5, 0.05)
    x = torch.tensor(rng.standard_normal((N, S, S, S, 1)), dtype=torch.float32).clamp(-1, 1)
    masks = ON.make_vnet_masks(rng, N, filters, L)
    return shapes, init, x, masks


@pytest.mark.parametrize("S,filters,L,N", [(32, 16, 4, 1), (32, 16, 3, 2)])
def test_vnet_blocks_teacher_forced(cuda, S, filters, L, N):
    """Every stage of custom_vnet (vnet_model.py:199-264) against the fp32 oracle, each fed the CUDA path's own input to
    that stage (its previous tap), so the check is per stage at bf16 noise level (2e-2) instead of compounding through
    the un-normalised error growth of a randomly initialised V-Net (see test_vnet_whole_network)."""
    import torch.nn.functional as F
    from oracle import nets as ON
    from van_gan_b200.vnet_model import custom_vnet
    from van_gan_b200 import engine as E
    shapes, init, x, masks = _vnet_case(S, filters, L, N)
    P = ON.to_torch(init, requires_grad=False)
    net = custom_vnet((S, S, S, 1), use_batch_norm=False, upsample_mode='upsample', dropout=0.5, filters=filters, num_layers=L,
                      output_activation='tanh')
    net.load(init)
    taps = {}
    out = net.forward(E.Tape(enabled=False), E.Var(x.cuda()), training=True, masks=[m.cuda() for m in masks], taps=taps)
    T = {k: v.data.float().cpu() for k, v in taps.items()}
    pool = lambda t: F.max_pool3d(t.permute(0, 4, 1, 2, 3), 2).permute(0, 2, 3, 4, 1)
    prev = x
    for l in range(L):
        ref = ON._vnet_block(P, "enc%d" % l, prev, masks[l])
        assert rel_l2(T["enc%d" % l], ref) < 2e-2, "enc%d" % l
        prev = pool(T["enc%d" % l])
    ref = ON._vnet_block(P, "bridge", prev, masks[L])
    assert rel_l2(T["bridge"], ref) < 2e-2, "bridge"
    prev = T["bridge"]
    for l in reversed(range(L)):
        u = ON.conv3d(ON.upsample2(prev), P["dec%d.up.conv.w" % l], P["dec%d.up.conv.b" % l], padding="same")
        ref = ON._vnet_block(P, "dec%d" % l, torch.cat([_bf(u), T["enc%d" % l]], dim=-1))
        assert rel_l2(T["dec%d" % l], ref) < 2e-2, "dec%d" % l
        prev = T["dec%d" % l]
    ref = torch.tanh(ON.conv3d(prev, P["head.w"], P["head.b"], padding="same"))
    assert rel_l2(out.data, ref) < 2e-2, "head"


# Measured (B200, deterministic): output 1.9e-2, aggregate gradient 0.19, worst variable (bridge.c2.in.beta) 0.34 against the
# bf16-storage oracle -- the conditioning described in the docstring below, not wiring: every block of this network, teacher-forced
# with the oracle's own activations, is within 1.5e-2 (test_vnet_blocks_teacher_forced).  A missing or mis-scaled edge gives >= 1.0
# on the variables behind it, so both bounds still reject a wrong backward graph.
AGG_TOL, VAR_TOL = 0.25, 0.45


def test_vnet_whole_network(cuda):
    """End to end, output and weight gradients.  A randomly initialised V-Net (conv -> ReLU -> InstanceNorm chains without
    residual paths) amplifies ANY perturbation ~2.2x per block: the fp32 oracle itself moves 1.3e-2 when its input is
    perturbed by 1e-3 (L=2), and its bf16-storage emulation differs from it by 5.8e-2.  The end-to-end bound is therefore
    set against the bf16-storage oracle with that conditioning stated, while wiring and per-op numerics are pinned by the
    exact per-op tests above (gather_pad, maxpool, relu-input norm, conv) and the teacher-forced stage test."""
    from oracle import nets as ON
    from van_gan_b200.vnet_model import custom_vnet
    from van_gan_b200 import engine as E
    S, filters, L, N = 32, 16, 2, 1
    shapes, init, x, masks = _vnet_case(S, filters, L, N)
    gy = torch.tensor(np.random.default_rng(3).standard_normal((N, S, S, S, 1)), dtype=torch.float32)
    ON.Emu.on = True
    try:
        P = ON.to_torch(init)
        y = ON.vnet_forward(P, x, L, masks)
        g = torch.autograd.grad((y * gy).sum(), list(P.values()))
    finally:
        ON.Emu.on = False
    gref = OrderedDict_zip(P.keys(), g)
    net = custom_vnet((S, S, S, 1), use_batch_norm=False, upsample_mode='upsample', dropout=0.5, filters=filters, num_layers=L,
                      output_activation='tanh')
    net.load(init)
    tape = E.Tape()
    out = net.forward(tape, E.Var(x.cuda()), training=True, masks=[m.cuda() for m in masks])
    assert rel_l2(out.data, y.detach()) < 5e-2
    net.zero_grad()
    tape.backward([(out, gy.cuda())], net.trainable_variables)
    gg = net.export_grads()
    num = sum(float(((torch.tensor(gg[k]).double() - gref[k].double()) ** 2).sum()) for k in shapes)
    den = sum(float((gref[k].double() ** 2).sum()) for k in shapes)
    agg = (num / den) ** 0.5
    per = {k: rel_l2(torch.tensor(gg[k]), gref[k]) for k in shapes}
    worst = max(per, key=per.get)
    print("vnet whole network: output %.2e, gradient aggregate %.2e, worst variable %s %.2e"
          % (rel_l2(out.data, y.detach()), agg, worst, per[worst]))
    assert agg < AGG_TOL, agg
    # every variable receives a gradient of the right scale (a missing edge in the backward graph would give 0 or O(1) error)
    for k in shapes:
        assert per[k] < VAR_TOL, (k, per[k])


def OrderedDict_zip(keys, vals):
    from collections import OrderedDict
    return OrderedDict(zip(keys, vals))


def test_vangan_train_step_with_vnet_generator(cuda):
    """VanGan(gen_i2s='vnet') runs a full train step on the CUDA path (vangan.py:97-110); losses finite, weights move."""
    from bench import Args, synth_batch
    from van_gan_b200.vangan import VanGan
    S = 32
    I, Sg = synth_batch(1, S, 5)
    gan = VanGan(Args(S, 1, 1), gen_i2s='vnet', gen_s2i='resUnet')
    w0 = gan.gen_IS.w.clone()
    res = gan.train_step(torch.tensor(I), torch.tensor(Sg))
    assert all(np.isfinite(v) for v in res.values()), res
    assert float((gan.gen_IS.w - w0).abs().max()) > 0
