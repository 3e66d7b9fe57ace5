"""Step-2 loss sensitivity of the two-step Adam test to kernel-path toggles (run once per env configuration)."""
import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_train_step import _setup, Args
from oracle import losses as OL, nets as ON, step as OS
from van_gan_b200.vangan import VanGan
S, b = 32, 1
real_I, real_S, init, _ = _setup(S, b, 1, 5, perturb=0.0)
rng = np.random.default_rng(6)
gan = VanGan(Args(S, b, 1), gen_i2s='resUnet', gen_s2i='resUnet')
for k, net in gan.networks.items():
    net.load(init[k])
out = []
for it in range(2):
    rand = {k: ON.make_disc_rand(rng, b, S) for k in ("S_real", "S_fake", "I_real", "I_fake")}
    rand_d = {k: ([t.cuda() for t in nz], [m.cuda() for m in mk]) for k, (nz, mk) in rand.items()}
    res_k = gan.train_step(real_I, real_S, rand=rand_d)
    out.append({k: round(v, 4) for k, v in res_k.items()})
print(os.environ.get("CFG", ""), out[1])
if os.environ.get("ORACLE"):
    cfg = OL.make_cfg(b, 1)
    P = {k: ON.to_torch(v) for k, v in init.items()}
    opts = {k: OS.Adam(list(v.keys())) for k, v in init.items()}
    rng = np.random.default_rng(6)
    for it in range(2):
        rand = {k: ON.make_disc_rand(rng, b, S) for k in ("S_real", "S_fake", "I_real", "I_fake")}
        res_o, _ = OS.train_step_dp(cfg, P, opts, real_I, real_S, [rand])
    print("oracle", {k: round(float(v), 4) for k, v in res_o.items()})
