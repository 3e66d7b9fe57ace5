"""GPU parity of the 3D ResNet generator (generator.py:7-73 with VanGan's arguments, vangan.py:88-95) against the CPU
oracle: the upsample+pad pass, every stage teacher-forced (forward, input gradient, parameter gradients), the whole network, and a
VanGan step with both generators 'resnet' (eager and graph replay)."""
import numpy as np
import pytest
import torch

from _blocks import agg_rel, bf, oracle_block_grads, smooth_grad

pytestmark = pytest.mark.gpu

LAYER_TOL = 2e-2


def rel_l2(a, b):
    a = torch.as_tensor(a).double().flatten().cpu()
    b = torch.as_tensor(b).double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_param_shapes_and_count(cuda):
    from oracle import nets as ON
    from van_gan_b200.generator import resnet_param_shapes
    assert list(ON.resnet_param_shapes().items()) == list(resnet_param_shapes().items())
    assert sum(int(np.prod(s)) for s in resnet_param_shapes().values()) == 25176897      # 25.2 M parameters


def test_upsample_pad_fwd_bwd(cuda):
    import torch.nn.functional as F
    from oracle import nets as ON
    from van_gan_b200 import engine as E
    rng = np.random.default_rng(2)
    a = bf(torch.tensor(rng.standard_normal((2, 3, 4, 5, 16)), dtype=torch.float32))
    ar = a.clone().requires_grad_(True)
    y = F.pad(ON.upsample2(ar).permute(0, 4, 1, 2, 3), (1, 2, 1, 2, 1, 2)).permute(0, 2, 3, 4, 1)     # TF 'same' for k4: 1 before, 2 after
    g = bf(torch.tensor(rng.standard_normal(y.shape), dtype=torch.float32))
    y.backward(g)
    tape = E.Tape()
    av = E.Var(a.to(torch.bfloat16).cuda())
    out = E.upsample_pad(tape, av, 1, 2)
    assert torch.equal(out.data.float().cpu(), y.detach())
    tape.backward([(out, g.to(torch.bfloat16).cuda())], [], wrt_vars=[av])
    assert rel_l2(av.grad.float(), ar.grad) < 4e-3            # bf16 store of an 8-term sum


def _case(S, filters, nr, N, seed=1):
    from oracle import nets as ON
    rng = np.random.default_rng(seed)
    shapes = ON.resnet_param_shapes(filters, 3, nr, 3)
    init = ON.init_params(shapes, 5, 0.05)
    x = torch.tensor(rng.standard_normal((N, S, S, S, 1)), dtype=torch.float32).clamp(-1, 1)
    masks = ON.make_resnet_masks(rng, N, filters, 3)
    return rng, shapes, init, x, masks


@pytest.mark.parametrize("S,filters,nr,N", [(32, 32, 2, 2)])
def test_resnet_stages_teacher_forced(cuda, S, filters, nr, N):
    """Every stage fed the CUDA path's own input to that stage and a given upstream gradient: output at 2e-2 of the fp32 oracle; input
    and parameter gradients at 2e-2 of the oracle with bf16 storage emulation (same arithmetic, same storage points), whose own distance
    from the fp32 oracle is printed and must stay small."""
    from oracle import nets as ON
    from van_gan_b200 import engine as E
    from van_gan_b200._lib import ACT_NONE, ACT_RELU, PAD_REFLECT, PAD_ZERO
    from van_gan_b200.generator import get_resnet_generator
    rng, shapes, init, x, masks = _case(S, filters, nr, N)
    P = ON.to_torch(init)
    net = get_resnet_generator((S, S, S, 1), filters=filters, num_downsampling_blocks=3, num_residual_blocks=nr, num_upsample_blocks=3,
                               name="g")
    net.load(init)
    taps = {}
    out = net.forward(E.Tape(enabled=False), E.Var(x.cuda()), training=True, masks=[m.cuda() for m in masks], taps=taps)
    # taps hold the tensors as the NEXT stage receives them (padded where the producer pads): strip the padding for the oracle
    def unpad(t, lo, hi):
        return t[:, lo:t.shape[1] - hi, lo:t.shape[2] - hi, lo:t.shape[3] - hi, :]
    T = {"c0": unpad(taps["c0"].data.float().cpu(), 1, 1), "down0": unpad(taps["down0"].data.float().cpu(), 1, 1),
         "down1": unpad(taps["down1"].data.float().cpu(), 1, 1), "down2": taps["down2"].data.float().cpu()}
    for j in range(nr):
        T["res%d" % j] = taps["res%d" % j].data.float().cpu()
    T["up0"], T["up1"] = taps["up0"].data.float().cpu(), taps["up1"].data.float().cpu()
    T["up2"] = unpad(taps["up2"].data.float().cpu(), 3, 3)
    report = []

    def check(name, fn, xin, pn, cuda_fn, fp32_in=False, cuda_in=None, crop=0):
        with torch.no_grad():
            yshape = tuple(fn(P, xin).shape)
        g = smooth_grad(rng, yshape)
        y, gx, gp = oracle_block_grads(fn, P, [xin], g, pn)
        ON.Emu.on = True
        try:
            _, gx_e, gp_e = oracle_block_grads(fn, P, [xin], g, pn)
        finally:
            ON.Emu.on = False
        tape = E.Tape()
        xc = xin if cuda_in is None else cuda_in
        xv = E.Var(xc.cuda() if fp32_in else xc.to(torch.bfloat16).cuda())
        o = cuda_fn(tape, xv)
        net.zero_grad()
        gdev = g.cuda() if o.data.dtype == torch.float32 else g.to(torch.bfloat16).cuda()
        tape.backward([(o, gdev)], net.trainable_variables, wrt_vars=[xv])
        gk = net.export_grads()
        gxk = xv.grad.float()
        if crop:
            gxk = gxk[:, crop:-crop, crop:-crop, crop:-crop, :]
        report.append((name, rel_l2(o.data.float(), y), rel_l2(gx_e[0], gx[0]), agg_rel(gp_e, gp, pn), rel_l2(gxk, gx_e[0]),
                       agg_rel(gk, gp_e, pn)))
        tape.clear()

    m0 = masks[0]
    # stage c0: compare the un-padded part of the producer's padded output
    def cuda_c0(tape, xv):
        h = net.conv0(tape, E.pad_noise(tape, xv))
        return net.norm0(tape, h, act=ACT_RELU, drop=m0.reshape(-1).cuda(), pad=(0, 0, PAD_ZERO))
    check("c0", lambda p, t: ON.resnet_stage_c0(p, t, m0), x, ["c0.conv.w", "c0.in.gamma", "c0.in.beta"], cuda_c0, fp32_in=True)
    prev = T["c0"]
    for i in range(3):
        conv, norm = net.down[i]
        mk = masks[i + 1]

        def cuda_down(tape, xv, conv=conv, norm=norm, mk=mk):
            p = E.gather_pad(tape, None, xv, up=1, pad=1, mode=PAD_REFLECT)
            return norm(tape, conv(tape, p), act=ACT_RELU, drop=mk.reshape(-1).cuda(), pad=(0, 0, PAD_ZERO))
        check("down%d" % i, lambda p, t, i=i, mk=mk: ON.resnet_stage_down(p, i, t, mk), prev,
              ["down%d.conv.w" % i, "down%d.in.gamma" % i, "down%d.in.beta" % i], cuda_down)
        prev = T["down%d" % i]
    for j in range(nr):
        c1, n1, c2, n2 = net.res[j]

        def cuda_res(tape, xv, c1=c1, n1=n1, c2=c2, n2=n2):
            p = E.gather_pad(tape, None, xv, up=1, pad=1, mode=PAD_REFLECT)
            c = n1(tape, c1(tape, p), act=ACT_RELU, pad=(1, 1, PAD_REFLECT))
            return n2(tape, c2(tape, c), act=ACT_NONE, residual=xv)
        check("res%d" % j, lambda p, t, j=j: ON.resnet_stage_res(p, j, t), prev, [n for n in P if n.startswith("res%d." % j)], cuda_res)
        prev = T["res%d" % j]
    for i in range(3):
        conv, norm = net.up[i]

        def cuda_up(tape, xv, conv=conv, norm=norm):
            return norm(tape, conv(tape, E.upsample_pad(tape, xv, 1, 2)), act=ACT_RELU, pad=(0, 0, PAD_ZERO))
        check("up%d" % i, lambda p, t, i=i: ON.resnet_stage_up(p, i, t), prev, ["up%d.conv.w" % i, "up%d.in.gamma" % i, "up%d.in.beta" % i],
              cuda_up)
        prev = T["up%d" % i]

    import torch.nn.functional as F
    prev_pad = F.pad(prev.permute(0, 4, 1, 2, 3), (3,) * 6).permute(0, 2, 3, 4, 1).contiguous()     # the zeros of the k7 'same' head
    check("out", lambda p, t: ON.resnet_stage_out(p, t), prev, ["out.conv.w", "out.conv.b"], lambda tape, xv: net.out(tape, xv),
          cuda_in=prev_pad, crop=3)
    for r in report:
        print("resnet stage %-6s fwd %.2e | bf16-storage oracle vs fp32 oracle: dx %.2e params %.2e | CUDA vs bf16-storage oracle: dx %.2e "
              "params %.2e" % r)
    for name, e_f, f_x, f_p, v_x, v_p in report:
        assert e_f < 2e-2, (name, e_f)
        assert f_x < 0.15 and f_p < 0.15, (name, f_x, f_p)
        assert v_x < LAYER_TOL and v_p < LAYER_TOL, (name, v_x, v_p)
    with torch.no_grad():
        ref = ON.resnet_stage_out(P, prev)
    assert rel_l2(out.data, ref) < 2e-2, "head"


def test_resnet_whole_network_forward(cuda):
    from oracle import nets as ON
    from van_gan_b200.generator import get_resnet_generator
    S, filters, nr, N = 32, 32, 2, 1
    rng, shapes, init, x, masks = _case(S, filters, nr, N, seed=4)
    P = ON.to_torch(init, requires_grad=False)
    ON.Emu.on = True
    try:
        y = ON.resnet_forward(P, x, 3, nr, 3, None)
    finally:
        ON.Emu.on = False
    net = get_resnet_generator((S, S, S, 1), filters=filters, num_downsampling_blocks=3, num_residual_blocks=nr, num_upsample_blocks=3,
                               name="g")
    net.load(init)
    out = net(x.numpy(), training=False)
    e = rel_l2(out, y)
    print("resnet whole network (inference) vs bf16-storage oracle: %.2e" % e)
    assert tuple(out.shape) == (N, S, S, S, 1) and e < 3e-2, e


def test_vangan_train_step_with_resnet_generators(cuda):
    """VanGan's DEFAULT generators (gen_i2s = gen_s2i = 'resnet', vangan.py:29-30): eager steps, graph capture, replay."""
    from bench import Args, synth_batch
    from van_gan_b200.vangan import VanGan
    S = 32
    I, Sg = synth_batch(1, S, 5)
    gan = VanGan(Args(S, 1, 1), gen_i2s='resnet', gen_s2i='resnet')
    assert sum(p.size for p in gan.gen_IS.trainable_variables) == 25176897
    w0 = gan.gen_SI.w.clone()
    for it in range(4):
        res = gan.train_step(torch.tensor(I).cuda(), torch.tensor(Sg).cuda())
        assert all(np.isfinite(v) for v in res.values()), (it, res)
    assert gan._graph is not None, "the graph path did not engage"
    assert float((gan.gen_SI.w - w0).abs().max()) > 0
    assert float(gan.gen_IS.params["c0.conv.w"].grad.abs().max()) > 0 and float(gan.gen_IS.params["out.conv.w"].grad.abs().max()) > 0
    res_t = gan.test_step(torch.tensor(I).cuda(), torch.tensor(Sg).cuda())
    assert all(np.isfinite(v) for v in res_t.values()), res_t
