import sys, os, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_train_step import Args, synth, rel_l2
from oracle import losses as OL, nets as ON, step as OS
from van_gan_b200.vangan import VanGan
from van_gan_b200 import engine as E

def agg(gk, go):
    num = sum(float(((torch.tensor(gk[n]).double() - go[n].double()) ** 2).sum()) for n in gk)
    den = sum(float((go[n].double() ** 2).sum()) for n in gk)
    return (num / den) ** 0.5

def run(S, b, emu, perturb):
    nd = 2; G = b * nd
    rng = np.random.default_rng(101)
    real_I, real_S = synth(rng, b, S)
    init = {"gen_IS": ON.init_params(ON.resunet_param_shapes(), 1, perturb), "gen_SI": ON.init_params(ON.resunet_param_shapes(), 2, perturb),
            "disc_I": ON.init_params(ON.disc_param_shapes(), 3, perturb), "disc_S": ON.init_params(ON.disc_param_shapes(), 4, perturb)}
    rand = {k: ON.make_disc_rand(rng, b, S) for k in ("S_real", "S_fake", "I_real", "I_fake")}
    cfg = OL.make_cfg(G, nd)
    P = {k: ON.to_torch(v) for k, v in init.items()}
    ON.Emu.on = emu
    t0 = time.time()
    res_o, grads_o, aux_o = OS.replica_grads(cfg, P, real_I, real_S, rand)
    ON.Emu.on = False
    print("oracle time", time.time() - t0)
    gan = VanGan(Args(S, G, nd), gen_i2s='resUnet', gen_s2i='resUnet'); gan.keep_last = True
    for k, net in gan.networks.items():
        net.load(init[k])
    rand_d = {k: ([t.cuda() for t in nz], [m.cuda() for m in mk]) for k, (nz, mk) in rand.items()}
    res_k = gan.train_step(real_I, real_S, rand=rand_d, apply=False)
    print("== S=%d b=%d emu=%s" % (S, b, emu))
    for k in ("fake_S", "fake_I", "cycled_S", "cycled_I", "disc_fake_S", "disc_real_S"):
        print("  fwd", k, "%.5f" % rel_l2(gan.last[k].data.float().cpu(), aux_o[k].detach()))
    for k in OS.RESULT_KEYS:
        print("  loss %-24s %.6f %.6f  rel %.2e" % (k, res_k[k], float(res_o[k]), abs(res_k[k] - float(res_o[k])) / abs(float(res_o[k]))))
    for name, net in gan.networks.items():
        print("  grad", name, "%.5f" % agg(net.export_grads(), grads_o[name]))

run(32, 1, True, 0.05)
run(32, 2, True, 0.05)
run(64, 1, True, 0.05)
run(64, 1, False, 0.05)
