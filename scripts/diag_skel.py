"""Diagnostic: repeat the soft-skeleton forward + backward on identical inputs and count deviating results."""
import os, sys, numpy as np, torch
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
from van_gan_b200 import clDice_func as K
R = int(sys.argv[1]) if len(sys.argv) > 1 else 100
shape = tuple(int(a) for a in sys.argv[2].split("x")) if len(sys.argv) > 2 else (2, 32, 32, 32)
kind = sys.argv[3] if len(sys.argv) > 3 else "tanh"
rng = np.random.default_rng(5)
raw = rng.standard_normal(shape + (1,)).astype(np.float32)
if kind == "tanh":       # generator-like output: smooth field pushed through a saturating tanh, mapped to [0, 1]
    from scipy.ndimage import gaussian_filter
    sm = gaussian_filter(raw[..., 0], sigma=(0, 2, 2, 2)) * 40.0
    x = (np.tanh(sm)[..., None] + 1.0) * 0.5
elif kind == "quant":
    x = np.round(rng.random(shape + (1,)) * 4) / 4
else:
    x = rng.random(shape + (1,))
x = torch.tensor(x.astype(np.float32)).cuda()
g = torch.tensor(rng.standard_normal(shape + (1,)).astype(np.float32)).cuda()
ref_s = ref_d = None
bad_s = bad_d = 0
worst = 0.0
for r in range(R):
    skel, bwd = K.soft_skel_with_grad(x, 15)
    skel = skel.clone()
    dx = bwd(g).clone()
    torch.cuda.synchronize()
    if ref_s is None:
        ref_s, ref_d = skel, dx
        continue
    if not torch.equal(skel, ref_s):
        bad_s += 1
    rel = float((dx - ref_d).double().norm() / ref_d.double().norm())
    worst = max(worst, rel)
    if rel > 1e-5:
        bad_d += 1
        if bad_d <= 3:
            diff = (dx - ref_d).abs().reshape(shape)
            idx = torch.nonzero(diff > 1e-4 * float(ref_d.abs().max()))
            print("  run %d: rel %.2e, %d voxels differ, first few %s" % (r, rel, idx.shape[0], idx[:6].tolist()))
print("%s %s env VG_SKEL_BWD=%s: %d runs, forward deviations %d, backward deviations (>1e-5) %d, worst rel %.2e, |dx| %.3e"
      % (kind, shape, os.environ.get("VG_SKEL_BWD"), R, bad_s, bad_d, worst, float(ref_d.double().norm())))
