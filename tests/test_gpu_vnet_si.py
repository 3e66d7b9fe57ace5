"""GPU parity of the V-Net gen_SI variant (vnet_model.py:80-146,149-268 with the arguments at vangan.py:135-149:
BatchNormalization, Conv3DTranspose k2 s2, filters 16) against the CPU oracle: the two new layers on their own at the per-layer
bound, every block teacher-forced (forward, input gradient, parameter gradients), inference with the moving statistics, the VanGan
step with both generators 'vnet', and the checkpoint round trip of the non-trainable variables."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from _blocks import agg_rel, bf, oracle_block_grads, smooth_grad

pytestmark = pytest.mark.gpu

LAYER_TOL = 2e-2
BLOCK_TOL_DX, BLOCK_TOL_P = 8e-2, 5e-2      # two convolutions deep (tests/test_gpu_parity_r2.py states where these come from)


def rel_l2(a, b):
    a = torch.as_tensor(a).double().flatten().cpu()
    b = torch.as_tensor(b).double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("C,S,N,pad,dt,tol", [(16, 10, 2, (1, 1, 1), torch.float32, 2e-5), (32, 8, 3, (0, 0, 0), torch.float32, 2e-5),
                                               (16, 10, 2, (1, 1, 1), torch.bfloat16, 2e-2)])
def test_batchnorm_on_relu_input(cuda, C, S, N, pad, dt, tol):
    """Conv3D(relu) -> BatchNormalization -> SpatialDropout3D -> ReflectionPadding3D (vnet_model.py:116-132, use_batch_norm=True):
    batch statistics over N*D*H*W, moving averages (momentum 0.99), backward with whole-batch reductions; then inference mode."""
    from oracle import nets as ON
    from van_gan_b200 import engine as E
    rng = np.random.default_rng(31)
    x = torch.tensor(rng.standard_normal((N, S, S + 1, S + 2, C)) * 1.3 + 0.2, dtype=torch.float32)
    x[1] = x[1] * 1.7 - 0.4                       # samples differ in scale: instance statistics would NOT match batch statistics
    gamma = torch.tensor(1 + 0.2 * rng.standard_normal(C), dtype=torch.float32)
    beta = torch.tensor(0.2 * rng.standard_normal(C), dtype=torch.float32)
    drop = torch.tensor((rng.random((N, 1, 1, 1, C)) > 0.5) / 0.5, dtype=torch.float32)
    if dt == torch.bfloat16:
        x = bf(x)
    state = OrderedDict([("n.moving_mean", torch.zeros(C)), ("n.moving_variance", torch.ones(C))])
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = ON.batch_norm(torch.relu(xr), gr, br, state, "n", training=True) * drop
    if pad[0]:
        y = ON.reflect_pad(y)
    gout = torch.tensor(rng.standard_normal(y.shape), dtype=torch.float32)
    if dt == torch.bfloat16:
        gout = bf(gout)
    y.backward(gout)
    net = E.Network("t", OrderedDict([("n.gamma", (C,)), ("n.beta", (C,))]))
    net.load({"n.gamma": gamma.numpy(), "n.beta": beta.numpy()})
    layer = E.BatchNorm(net, "n", C)
    tape = E.Tape()
    xv = E.Var(x.to(dt).cuda())
    out = layer(tape, xv, training=True, drop=drop.reshape(-1).cuda(), pad=pad, relu_input=True)
    assert rel_l2(out.data.float(), y.detach()) < (1e-5 if dt == torch.float32 else 1.5e-2)
    assert rel_l2(net.buffers["n.moving_mean"], state["n.moving_mean"]) < 1e-5
    assert rel_l2(net.buffers["n.moving_variance"], state["n.moving_variance"]) < 1e-5
    tape.backward([(out, gout.to(dt).cuda())], net.trainable_variables, wrt_vars=[xv])
    e = (rel_l2(xv.grad.float(), xr.grad), rel_l2(net.params["n.gamma"].grad, gr.grad), rel_l2(net.params["n.beta"].grad, br.grad))
    print("batchnorm C%d N%d %s: dx %.2e dgamma %.2e dbeta %.2e" % (C, N, dt, *e))
    assert max(e) < tol, e
    # inference: the moving statistics (after one update), no dropout
    yi = ON.batch_norm(torch.relu(x), gamma, beta, state, "n", training=False)
    oi = layer(E.Tape(enabled=False), E.Var(x.to(dt).cuda()), training=False, relu_input=True)
    assert rel_l2(oi.data.float(), yi) < (1e-5 if dt == torch.float32 else 1.5e-2)
    # a second inference call must not move the statistics
    assert rel_l2(net.buffers["n.moving_mean"], state["n.moving_mean"]) < 1e-5


@pytest.mark.parametrize("Cin,Cout,S,N", [(32, 16, 6, 2), (64, 32, 5, 1), (256, 128, 4, 2)])
def test_conv3d_transpose_k2s2(cuda, Cin, Cout, S, N):
    """Conv3DTranspose(filters, (2,2,2), strides 2, 'same') (vnet_model.py:245), kernel in Keras layout (2,2,2,Cout,Cin): forward,
    input gradient, kernel and bias gradients against the oracle on bf16-exact operands."""
    from oracle import nets as ON
    from van_gan_b200 import engine as E
    rng = np.random.default_rng(41)
    x = bf(torch.tensor(rng.standard_normal((N, S, S + 1, S + 2, Cin)), dtype=torch.float32))
    w = bf(torch.tensor(rng.standard_normal((2, 2, 2, Cout, Cin)) / np.sqrt(Cin), dtype=torch.float32))
    b = torch.tensor(0.1 * rng.standard_normal(Cout), dtype=torch.float32)
    xr, wr, br_ = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = ON.conv3d_transpose_k2s2(xr, wr, br_)
    assert tuple(y.shape) == (N, 2 * S, 2 * (S + 1), 2 * (S + 2), Cout)
    g = bf(torch.tensor(rng.standard_normal(y.shape), dtype=torch.float32))
    y.backward(g)
    net = E.Network("t", OrderedDict([("up.w", (2, 2, 2, Cout, Cin)), ("up.b", (Cout,))]))
    layer = E.Conv3DTranspose(net, "up", Cin, Cout)
    net.load({"up.w": w.numpy(), "up.b": b.numpy()})
    tape = E.Tape()
    xv = E.Var(x.to(torch.bfloat16).cuda())
    out = layer(tape, xv)
    net.zero_grad()
    tape.backward([(out, g.to(torch.bfloat16).cuda())], net.trainable_variables, wrt_vars=[xv])
    e = dict(fwd=rel_l2(out.data.float(), y.detach()), dx=rel_l2(xv.grad.float(), xr.grad), dw=rel_l2(net.params["up.w"].grad, wr.grad),
             db=rel_l2(net.params["up.b"].grad, br_.grad))
    print("conv3d_transpose %d->%d: %s" % (Cin, Cout, {k: "%.2e" % v for k, v in e.items()}))
    assert e["fwd"] < 4e-3 and e["dx"] < 4e-3, e          # bf16 store of the output / of dx
    assert e["dw"] < 1e-4 and e["db"] < 1e-4, e           # fp32 accumulation of bf16-exact products
    # a second backward must ADD to the gradient buffers (the persistent tape runs several sweeps)
    tape2 = E.Tape()
    out2 = layer(tape2, E.Var(x.to(torch.bfloat16).cuda()))
    tape2.backward([(out2, g.to(torch.bfloat16).cuda())], net.trainable_variables)
    assert rel_l2(net.params["up.w"].grad, 2 * wr.grad) < 1e-4


def _si_case(S, filters, L, N, seed=1, si=True):
    from oracle import nets as ON
    rng = np.random.default_rng(seed)
    shapes = ON.vnet_param_shapes(filters, L, 1, use_batch_norm=si, deconv=si)
    init = ON.init_params(shapes, 5, 0.05)
    x = torch.tensor(rng.standard_normal((N, S, S, S, 1)), dtype=torch.float32).clamp(-1, 1)
    masks = ON.make_vnet_masks(rng, N, filters, L)
    return rng, shapes, init, x, masks


def test_param_shapes_match_oracle(cuda):
    from oracle import nets as ON
    from van_gan_b200.vnet_model import vnet_param_shapes
    for bn, dec in ((False, False), (True, True), (True, False), (False, True)):
        assert list(ON.vnet_param_shapes(16, 4, 1, bn, dec).items()) == \
            list(vnet_param_shapes(16, 4, 1, bn, 'deconv' if dec else 'upsample').items())
    n = sum(int(np.prod(s)) for s in vnet_param_shapes(16, 4, 1, True, 'deconv').values())
    assert n == 5646385, n                      # trainable parameters of the gen_SI variant
    assert sum(int(np.prod(s)) for s in vnet_param_shapes(32, 4, 1, False, 'upsample').values()) == 25888737   # gen_IS variant


@pytest.mark.parametrize("S,filters,L,N,si", [(32, 16, 4, 2, True), (32, 16, 4, 2, False)])
def test_vnet_si_blocks_teacher_forced(cuda, S, filters, L, N, si):
    """Every stage of custom_vnet -- si=True: the gen_SI variant (use_batch_norm=True, upsample_mode='deconv'); si=False: the gen_IS
    variant (InstanceNormalization, UpSampling3D + Conv3D) -- fed the CUDA path's own input to that stage: forward at 2e-2, and, with a
    given upstream gradient, the input gradient and the stage's parameter gradients against the fp32 oracle and against the oracle with
    bf16 storage emulation (the per-stage gradient evidence the whole-network test of tests/test_gpu_vnet.py cannot give)."""
    import torch.nn.functional as F
    from oracle import nets as ON
    from van_gan_b200.vnet_model import custom_vnet
    from van_gan_b200 import engine as E
    from van_gan_b200._lib import PAD_REFLECT
    rng, shapes, init, x, masks = _si_case(S, filters, L, N, si=si)
    P = ON.to_torch(init)
    net = custom_vnet((S, S, S, 1), use_batch_norm=si, upsample_mode='deconv' if si else 'upsample', dropout=0.5, filters=filters,
                      num_layers=L, output_activation='tanh')
    net.load(init)
    taps = {}
    out = net.forward(E.Tape(enabled=False), E.Var(x.cuda()), training=True, masks=[m.cuda() for m in masks], taps=taps)
    T = {k: v.data.float().cpu() for k, v in taps.items()}
    moving = {k: v.clone() for k, v in net.buffers.items()}          # after exactly ONE training forward
    pool = lambda t: F.max_pool3d(t.permute(0, 4, 1, 2, 3), 2).permute(0, 2, 3, 4, 1)
    state = ON.vnet_bn_state(filters, L)
    report = []

    def check_block(name, blk, xin, mask, is_input=False):
        pn = [n for n in P if n.startswith(name + ".c")]
        fn = lambda p, t: ON._vnet_block(p, name, t, mask, state, True)
        with torch.no_grad():
            yshape = tuple(fn(P, xin).shape)
        g = smooth_grad(rng, yshape)
        y, gx, gp = oracle_block_grads(fn, P, [xin], g, pn)
        ON.Emu.on = True                              # the oracle's own arithmetic with bf16 storage at the CUDA path's storage points
        try:
            _, gx_e, gp_e = oracle_block_grads(fn, P, [xin], g, pn)
        finally:
            ON.Emu.on = False
        floor = (rel_l2(gx_e[0], gx[0]), agg_rel(gp_e, gp, pn))
        emu = (gx_e[0], gp_e)
        tape = E.Tape()
        xv = E.Var(xin.cuda() if is_input else xin.to(torch.bfloat16).cuda())
        xp = E.pad_noise(tape, xv) if is_input else E.gather_pad(tape, None, xv, up=1, pad=1, mode=PAD_REFLECT)
        o = blk(tape, xp, drop=None if mask is None else mask.reshape(-1).cuda(), training=True)
        net.zero_grad()
        tape.backward([(o, g.to(torch.bfloat16).cuda())], net.trainable_variables, wrt_vars=[xv])
        gk = net.export_grads()
        report.append((name, rel_l2(o.data.float(), y), rel_l2(xv.grad.float(), gx[0]), agg_rel(gk, gp, pn), floor[0], floor[1],
                       rel_l2(xv.grad.float(), emu[0]), agg_rel(gk, emu[1], pn)))
        tape.clear()

    prev = x
    for l in range(L):
        check_block("enc%d" % l, net.enc[l], prev, masks[l], is_input=(l == 0))
        prev = pool(T["enc%d" % l])
    check_block("bridge", net.bridge, prev, masks[L])
    prev = T["bridge"]
    for l in reversed(range(L)):
        with torch.no_grad():
            if si:
                u = ON.conv3d_transpose_k2s2(prev, P["dec%d.up.w" % l], P["dec%d.up.b" % l])
            else:
                u = ON.conv3d(ON.upsample2(prev), P["dec%d.up.conv.w" % l], P["dec%d.up.conv.b" % l], padding="same")
        check_block("dec%d" % l, net.dec[l], torch.cat([bf(u), T["enc%d" % l]], dim=-1), None)
        prev = T["dec%d" % l]
    for r in report:
        print("vnet %s block %-7s fwd %.2e  dx %.2e  params %.2e   (bf16-storage oracle vs fp32 oracle: dx %.2e params %.2e;"
              " CUDA vs bf16-storage oracle: dx %.2e params %.2e)" % (("gen_SI" if si else "gen_IS",) + r))
    for name, e_f, e_x, e_p, f_x, f_p, v_x, v_p in report:
        assert v_x < LAYER_TOL and v_p < LAYER_TOL, (name, v_x, v_p)      # same arithmetic, same storage points: north_star's 2e-2
        # bound: the per-block figures of tests/test_gpu_parity_r2.py, or -- for the small deep stages, where a block holds only a few
        # thousand ReLU units and bf16 storage alone moves the gradient more -- 1.25 x what the oracle's own bf16-storage emulation
        # moves + 2e-2; the emulation floor itself must stay small for the bound to mean anything
        assert f_x < 0.15 and f_p < 0.15, (name, f_x, f_p)
        assert e_f < 2e-2, (name, e_f)
        assert e_x < max(BLOCK_TOL_DX, 1.25 * f_x + 2e-2), (name, e_x, f_x)
        assert e_p < max(BLOCK_TOL_P, 1.25 * f_p + 2e-2), (name, e_p, f_p)
    with torch.no_grad():
        ref = torch.tanh(ON.conv3d(prev, P["head.w"], P["head.b"], padding="same"))
    assert rel_l2(out.data, ref) < 2e-2, "head"
    if not si:
        return
    # the CUDA network's moving statistics after ONE training forward == the oracle's after one whole-network training forward
    state2 = ON.vnet_bn_state(filters, L)
    with torch.no_grad():
        ON.vnet_forward(P, x, L, masks, bn_state=state2, training=True)
    worst = max(rel_l2(moving[k], v) for k, v in state2.items() if k.startswith("enc0"))
    assert worst < 2e-2, worst          # first block: same inputs on both sides; deeper ones compound the bf16 forward error


def test_vnet_si_inference_uses_moving_statistics(cuda):
    from oracle import nets as ON
    from van_gan_b200.vnet_model import custom_vnet
    S, filters, L, N = 32, 16, 2, 1
    rng, shapes, init, x, masks = _si_case(S, filters, L, N, seed=3)
    P = ON.to_torch(init, requires_grad=False)
    net = custom_vnet((S, S, S, 1), use_batch_norm=True, upsample_mode='deconv', dropout=0.5, filters=filters, num_layers=L,
                      output_activation='tanh')
    net.load(init)
    state = ON.vnet_bn_state(filters, L)
    for k in state:                                   # non-trivial statistics on both sides
        state[k] = torch.tensor(np.abs(rng.standard_normal(state[k].shape)) * 0.5 + 0.5, dtype=torch.float32)
        net.buffers[k].copy_(state[k].cuda())
    y = ON.vnet_forward(P, x, L, None, bn_state=state, training=False)
    out = net(x.numpy(), training=False)
    e = rel_l2(out, y)
    print("vnet gen_SI inference: %.2e" % e)
    assert e < 3e-2, e


def test_vangan_train_step_with_both_vnet_generators(cuda, tmp_path):
    """VanGan(gen_i2s='vnet', gen_s2i='vnet') (vangan.py:97-110,135-149): a full train step on the CUDA path, the inference-mode
    test_step, and the checkpoint round trip of the BatchNormalization moving statistics."""
    from bench import Args, synth_batch
    from van_gan_b200.vangan import VanGan
    S = 32
    I, Sg = synth_batch(2, S, 5)
    args = Args(S, 2, 1)
    args.output_dir = str(tmp_path)
    gan = VanGan(args, gen_i2s='vnet', gen_s2i='vnet')
    assert gan.gen_SI.use_batch_norm and gan.gen_SI.upsample_mode == 'deconv' and gan.gen_SI.filters == 16
    w0 = gan.gen_SI.w.clone()
    mm0 = {k: v.clone() for k, v in gan.gen_SI.buffers.items()}
    res = gan.train_step(torch.tensor(I), torch.tensor(Sg))
    assert all(np.isfinite(v) for v in res.values()), res
    assert float((gan.gen_SI.w - w0).abs().max()) > 0
    up = gan.gen_SI.params["dec0.up.w"]
    assert float(up.grad.abs().max()) > 0 and float(gan.gen_SI.params["enc0.c1.bn.gamma"].grad.abs().max()) > 0
    assert any(float((gan.gen_SI.buffers[k] - mm0[k]).abs().max()) > 0 for k in mm0)
    res_t = gan.test_step(torch.tensor(I), torch.tensor(Sg))
    assert all(np.isfinite(v) for v in res_t.values()), res_t
    gan.save_checkpoint(0)
    other = VanGan(args, gen_i2s='vnet', gen_s2i='vnet', seed=99)
    assert other.load_checkpoint(epoch=1)
    for k, v in gan.gen_SI.buffers.items():
        assert torch.equal(other.gen_SI.buffers[k], v), k
    assert torch.equal(other.gen_SI.w, gan.gen_SI.w)
    res_o = other.test_step(torch.tensor(I), torch.tensor(Sg))
    for k in res_t:
        assert abs(res_o[k] - res_t[k]) <= 1e-5 * abs(res_t[k]) + 1e-7, (k, res_o[k], res_t[k])
