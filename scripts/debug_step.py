import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_train_step import Args, synth, rel_l2
from oracle import losses as OL, nets as ON, step as OS
from van_gan_b200.vangan import VanGan
from van_gan_b200 import engine as E
S, b, nd = 32, 1, 2
G = b * nd
rng = np.random.default_rng(101)
real_I, real_S = synth(rng, b, S)
init = {"gen_IS": ON.init_params(ON.resunet_param_shapes(), 1, 0.05), "gen_SI": ON.init_params(ON.resunet_param_shapes(), 2, 0.05),
        "disc_I": ON.init_params(ON.disc_param_shapes(), 3, 0.05), "disc_S": ON.init_params(ON.disc_param_shapes(), 4, 0.05)}
rand = {k: ON.make_disc_rand(rng, b, S) for k in ("S_real", "S_fake", "I_real", "I_fake")}
cfg = OL.make_cfg(G, nd)
P = {k: ON.to_torch(v) for k, v in init.items()}
res_o, grads_o, aux_o = OS.replica_grads(cfg, P, real_I, real_S, rand)
gan = VanGan(Args(S, G, nd), gen_i2s='resUnet', gen_s2i='resUnet'); gan.keep_last = True
for k, net in gan.networks.items():
    net.load(init[k])
rand_d = {k: ([t.cuda() for t in nz], [m.cuda() for m in mk]) for k, (nz, mk) in rand.items()}
result = {}
result, tI, tS, dI, dS, _, _ = gan.compute_losses(real_I, real_S, result, training=True, rand=rand_d)
for k in ("fake_S", "fake_I", "cycled_S", "cycled_I", "disc_fake_S", "disc_real_S"):
    print("fwd", k, rel_l2(gan.last[k].data.float().cpu(), aux_o[k].detach()))
# seeds
go = torch.autograd.grad(res_o["total_IS_loss"], [aux_o["cycled_S"], aux_o["disc_fake_S"], aux_o["fake_S"]], retain_graph=True)
seeds = tI.seeds()
for v, g in seeds:
    for k in gan.last:
        if gan.last[k] is v:
            print("seed for", k, tuple(g.shape), float(g.float().norm()))
gc = sum(g for v, g in seeds if v is gan.last["cycled_S"]).cpu()
gd = sum(g for v, g in seeds if v is gan.last["disc_fake_S"]).cpu()
print("seed cycled_S", rel_l2(gc, go[0]), "seed disc_fake_S", rel_l2(gd, go[1]))
net = gan.gen_IS
net.zero_grad()
gan.tape.backward(seeds, net.trainable_variables, wrt_vars=[gan.last["fake_S"]])
print("dL/dfake_S", rel_l2(gan.last["fake_S"].grad.cpu(), go[2]) if gan.last["fake_S"].grad is not None else None)
g = net.export_grads()
for n in g:
    print("%-22s %.4f  |ref| %.3e" % (n, rel_l2(torch.tensor(g[n]), grads_o["gen_IS"][n]), float(grads_o["gen_IS"][n].norm())))
