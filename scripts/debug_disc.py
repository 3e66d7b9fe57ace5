import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_train_step import synth, rel_l2
from oracle import nets as ON
from van_gan_b200 import engine as E
from van_gan_b200.discriminator import get_discriminator
S = 32
rng = np.random.default_rng(3)
real_I, real_S = synth(rng, 1, S)
initd = ON.init_params(ON.disc_param_shapes(), 3, 0.05)
nz, mk = ON.make_disc_rand(rng, 1, S)
d = get_discriminator((S, S, S, 1), filters=64, use_dropout=True, use_input_noise=True, use_layer_noise=True, name='d'); d.load(initd)
for emu in (False, True):
    Pd = ON.to_torch(initd)
    ON.Emu.on = emu
    xin = real_S.clone().requires_grad_(True)
    taps = {}
    yo = ON.disc_forward(Pd, xin, nz, mk, taps=taps)
    gu = torch.tensor(np.random.default_rng(9).standard_normal(yo.shape), dtype=torch.float32)
    for t in taps.values(): t.retain_grad()
    go = torch.autograd.grad(yo, list(Pd.values()) + [xin] , gu)
    ON.Emu.on = False
    tape = E.Tape(); xv = E.Var(real_S.cuda())
    out = d.forward(tape, xv, training=True, noise=[t.cuda() for t in nz], masks=[m.cuda() for m in mk])
    d.zero_grad()
    tape.backward([(out, gu.cuda())], d.trainable_variables, wrt_vars=[xv])
    gk = d.export_grads()
    print("emu", emu, "fwd %.5f dx %.5f" % (rel_l2(out.data.cpu(), yo.detach()), rel_l2(xv.grad.cpu(), go[-1])))
    for n, g in zip(Pd.keys(), go[:-1]):
        print("   %-14s %.5f |ref| %.3e" % (n, rel_l2(torch.tensor(gk[n]), g), float(g.norm())))
