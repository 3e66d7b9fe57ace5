"""Diagnostic: run the SAME eager train step repeatedly from one state and report which parameters' gradients deviate between runs
(the weight-gradient atomics give ~1e-7; anything larger is a race or an uninitialised read)."""
import os, sys, numpy as np, torch
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
from test_gpu_train_step import Args, synth
from van_gan_b200.vangan import VanGan
S, b, R = 32, 2, int(sys.argv[1]) if len(sys.argv) > 1 else 10
# record what the soft-skeleton backward sees and returns in every run
from van_gan_b200 import clDice_func as _K
TRACE = []
_orig = _K.soft_skel_with_grad
def _wrapped(img, iters):
    skel, bwd = _orig(img, iters)
    rec = {"img": img.clone(), "skel": skel.clone()}
    TRACE.append(rec)
    def bwd2(g):
        rec["g"] = g.clone()
        dx = bwd(g)
        rec["dx"] = dx.clone()
        if os.environ.get("VG_DIAG_RECHECK"):
            # the same call again, and the tile kernel on the same inputs
            rec["dx_again"] = bwd(g).clone()
            os.environ["VG_SKEL_BWD"] = "tile"
            rec["dx_tile"] = bwd(g).clone()
            del os.environ["VG_SKEL_BWD"]
        return dx
    return skel, bwd2
_K.soft_skel_with_grad = _wrapped
rng = np.random.default_rng(31)
batches = [synth(rng, b, S) for _ in range(4)]
gan = VanGan(Args(S, b, 1), gen_i2s='resUnet', gen_s2i='resUnet', seed=77)
gan.use_graph = False
for I, Sg in batches[:3]:
    gan.train_step(I.cuda(), Sg.cuda())
snap = {k: (net.w.clone(), net.m.clone(), net.v.clone(), net.step_count) for k, net in gan.networks.items()}
step0 = gan.step
runs = []
for r in range(R):
    for k, net in gan.networks.items():
        w, m, v, sc = snap[k]
        net.w.copy_(w); net.m.copy_(m); net.v.copy_(v); net.step_count = sc
        net.repack()
    gan.step = step0
    I, Sg = batches[3]
    res = gan.train_step(I.cuda(), Sg.cuda())
    torch.cuda.synchronize()
    runs.append((res, {k: net.g.clone() for k, net in gan.networks.items()}))
    runs[-1] += (TRACE[-1],)
# majority reference: the run whose gradients agree with most others
for k, net in gan.networks.items():
    G = [r_[1][k].double() for r_ in runs]
    d = [[float((G[i] - G[j]).norm() / G[i].norm()) for j in range(R)] for i in range(R)]
    ref = max(range(R), key=lambda i: sum(1 for j in range(R) if d[i][j] < 1e-5))
    odd = [i for i in range(R) if d[ref][i] >= 1e-5]
    print("%-7s reference run %d, deviating runs %s" % (k, ref, ["%d:%.1e" % (i, d[ref][i]) for i in odd]))
    for i in odd[:2]:
        for name, p in net.params.items():
            a, c = G[ref][p.offset:p.offset + p.size], G[i][p.offset:p.offset + p.size]
            rd = float((a - c).norm() / (a.norm() + 1e-30))
            if rd > 1e-5:
                print("     run %d  %-28s rel %.2e  |g| %.3e" % (i, name, rd, float(a.norm())))
print("losses of run 0:", {k: round(v, 6) for k, v in runs[0][0].items()})
for i in range(1, R):
    dl = {k: abs(runs[i][0][k] - runs[0][0][k]) for k in runs[0][0]}
    big = {k: "%.1e" % v for k, v in dl.items() if v > 1e-6 * abs(runs[0][0][k]) + 1e-9}
    if big:
        print("run %d loss deviations: %s" % (i, big))

t0 = runs[0][2]
for i in range(R):
    t = runs[i][2]
    line = "run %2d skeleton: " % i
    for key in ("img", "skel", "g", "dx"):
        d = float((t[key] - t0[key]).double().norm() / (t0[key].double().norm() + 1e-30))
        line += "%s %.1e  " % (key, d)
    if "dx_tile" in t:
        line += "| same call again %.1e, tile kernel %.1e" % (float((t["dx_again"] - t["dx"]).double().norm() / t["dx"].double().norm()),
                                                               float((t["dx_tile"] - t["dx"]).double().norm() / t["dx"].double().norm()))
    print(line)
