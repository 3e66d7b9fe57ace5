"""Generates tests/golden/step128_b1.npz: the fp32 CPU oracle's full train step at BASELINE's size (1 x 128^3 per replica,
N_DEVICES=1), reduced to what fits in a fixture: the ten losses, the per-variable gradient norms of the four networks and K
random-sign projections of each network's flat gradient (seeded), from which the GPU test estimates ||g_cuda - g_oracle|| / ||g_oracle||
and the cosine without shipping 165 MB of gradients.  TEST INFRASTRUCTURE; takes ~10 min and ~40 GB of host memory.

    python scripts/make_golden_128.py [S]      (S = 128 by default)
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

def main():
    from oracle import losses as OL, nets as ON, step as OS
    from _blocks import K_PROJ, projections
    from test_gpu_train_step import _setup
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    b, nd, seed = 1, 1, 128
    torch.set_num_threads(os.cpu_count())
    real_I, real_S, init, rand = _setup(S, b, nd, seed)
    cfg = OL.make_cfg(b * nd, nd)
    t = time.time()
    res, grads, _ = OS.replica_grads(cfg, {k: ON.to_torch(v) for k, v in init.items()}, real_I, real_S, rand)
    print("oracle step %.1f s" % (time.time() - t), flush=True)
    out = {"S": S, "seed": seed, "k_proj": K_PROJ}
    for k, v in res.items():
        out["loss/" + k] = float(v.detach())
    for i, (net, g) in enumerate(grads.items()):
        flat = torch.cat([g[n].detach().double().flatten() for n in g])
        out["norm/" + net] = float(flat.norm())
        out["proj/" + net] = projections(flat, 1000 + i)
        out["tnorm/" + net] = np.array([float(g[n].detach().double().norm()) for n in g])
    path = os.path.join(ROOT, "tests", "golden", "step%d_b1.npz" % S)
    np.savez(path, **out)
    print("wrote", path, {k: out[k] for k in out if k.startswith("loss/")})


if __name__ == "__main__":
    main()
