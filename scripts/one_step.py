"""ONE eager VanGan.train_step of bench.py's workload (8 x 128^3, same synthetic batch and seeds) -- the unit the ncu launch list
is taken from: under `ncu --metrics gpu__time_duration.sum` every launch costs ~55 ms, so the full bench command (7 steps,
14 400 launches) is ~13 min of GPU time for the same per-step list."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import Args, synth_batch
from van_gan_b200.vangan import VanGan
S, b = 128, (int(sys.argv[1]) if len(sys.argv) > 1 else 8)
I, Sg = synth_batch(b, S, 100)
gan = VanGan(Args(S, b, 1), gen_i2s='resUnet', gen_s2i='resUnet', seed=1234)
gan.use_graph = False
res = gan.train_step(torch.tensor(I).cuda(), torch.tensor(Sg).cuda())
torch.cuda.synchronize()
print({k: round(v, 4) for k, v in res.items()})
