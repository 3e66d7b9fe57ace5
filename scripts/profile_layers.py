"""Per-(ABI call, shape) device time inside one train step at b x S^3 (CUDA events around every ABI call)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import Args, synth_batch
from van_gan_b200 import _lib
from van_gan_b200.vangan import VanGan
S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
b = int(sys.argv[2]) if len(sys.argv) > 2 else 8
I, Sg = synth_batch(b, S, 3)
gan = VanGan(Args(S, b, 1), gen_i2s='resUnet', gen_s2i='resUnet')
gan.use_graph = False   # CUDA events around every ABI call need eager launches ...
gan._side = gan._wg_side = None   # ... on ONE stream (side streams would add other branches' kernels to every event pair)
dI, dS = torch.tensor(I).cuda(), torch.tensor(Sg).cuda()
for i in range(2):
    gan.train_step(dI, dS)
torch.cuda.synchronize()
prof = _lib.Profiler([], detail=True)
_lib.PROFILER = prof
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); gan.train_step(dI, dS); e1.record()
_lib.PROFILER = None
torch.cuda.synchronize()
tot = e0.elapsed_time(e1)
summ = prof.summary()
acc = sum(v["ms"] for v in summ.values())
print("step %.1f ms; sum of ABI calls %.1f ms" % (tot, acc))
for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"])[:int(os.environ.get('VG_TOP', '70'))]:
    tf = (" %7.1f TFLOP/s" % (v["work"] / v["ms"] / 1e9)) if v["work"] else ""
    print("%-58s n=%3d %8.3f ms %5.1f%%%s" % (k, v["calls"], v["ms"], 100 * v["ms"] / tot, tf))
