"""One warm train step + one profiled train step at 1 x S^3 (for `ncu` launch lists)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import Args, synth_batch
from van_gan_b200.vangan import VanGan
S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
b = int(sys.argv[2]) if len(sys.argv) > 2 else 1
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
I, Sg = synth_batch(b, S, 3)
gan = VanGan(Args(S, b, 1), gen_i2s='resUnet', gen_s2i='resUnet')
dI, dS = torch.tensor(I).cuda(), torch.tensor(Sg).cuda()
for i in range(steps):
    torch.cuda.synchronize()
    r = gan.train_step(dI, dS)
    torch.cuda.synchronize()
print(r)
from van_gan_b200 import _lib
print('kernel launches', _lib.lib().vg_launch_count(), 'tcgen05 conv launches', _lib.lib().vg_tc_launch_count())
