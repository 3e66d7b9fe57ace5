import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from scipy import ndimage
from test_gpu_train_step import synth, rel_l2
from oracle import nets as ON
from van_gan_b200 import engine as E
from van_gan_b200.resunet_model import ResUNet
from van_gan_b200.discriminator import get_discriminator
def agg(gk, go):
    num = sum(float(((torch.tensor(gk[n]).double() - go[n].double()) ** 2).sum()) for n in gk)
    den = sum(float((go[n].double() ** 2).sum()) for n in gk)
    return (num / den) ** 0.5
for S in (32, 64):
    rng = np.random.default_rng(3)
    real_I, real_S = synth(rng, 1, S)
    g_up = torch.tensor(ndimage.gaussian_filter(rng.standard_normal((1, S, S, S, 1)), (0, 1, 1, 1, 0)), dtype=torch.float32)
    init = ON.init_params(ON.resunet_param_shapes(), 1, 0.05)
    P = ON.to_torch(init)
    xin = real_I.clone().requires_grad_(True)
    yo = ON.resunet_forward(P, xin)
    go = torch.autograd.grad(yo, list(P.values()) + [xin], g_up)
    net = ResUNet((S, S, S, 1), upsample_mode='simple'); net.load(init)
    tape = E.Tape(); xv = E.Var(real_I.cuda())
    out = net.forward(tape, xv)
    net.zero_grad()
    tape.backward([(out, g_up.cuda())], net.trainable_variables)
    gk = net.export_grads()
    print("GEN S=%d fwd %.4f  grads %.4f" % (S, rel_l2(out.data.cpu(), yo.detach()), agg(gk, dict(zip(P.keys(), go[:-1])))))
    worst = sorted(((rel_l2(torch.tensor(gk[n]), g), n) for n, g in zip(P.keys(), go[:-1]) if float(g.norm()) > 1e-3), reverse=True)[:5]
    print("   worst", worst)
    # discriminator
    initd = ON.init_params(ON.disc_param_shapes(), 3, 0.05)
    Pd = ON.to_torch(initd)
    nz, mk = ON.make_disc_rand(rng, 1, S)
    xin = real_S.clone().requires_grad_(True)
    yo = ON.disc_forward(Pd, xin, nz, mk)
    gu = torch.tensor(rng.standard_normal(yo.shape), dtype=torch.float32)
    go = torch.autograd.grad(yo, list(Pd.values()) + [xin], gu)
    d = get_discriminator((S, S, S, 1), filters=64, use_dropout=True, use_input_noise=True, use_layer_noise=True, name='d'); d.load(initd)
    tape = E.Tape(); xv = E.Var(real_S.cuda())
    out = d.forward(tape, xv, training=True, noise=[t.cuda() for t in nz], masks=[m.cuda() for m in mk])
    d.zero_grad()
    tape.backward([(out, gu.cuda())], d.trainable_variables, wrt_vars=[xv])
    print("DISC S=%d fwd %.4f  grads %.4f  dx %.4f" % (S, rel_l2(out.data.cpu(), yo.detach()), agg(d.export_grads(), dict(zip(Pd.keys(), go[:-1]))), rel_l2(xv.grad.cpu(), go[-1])))
