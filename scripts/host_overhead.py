"""Host-side cost of one train step: wall time per step at tiny (host-bound) and real sizes, b=1."""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import Args, synth_batch
from van_gan_b200.vangan import VanGan
for S in (32, 64, 128):
    I, Sg = synth_batch(1, S, 3)
    gan = VanGan(Args(S, 1, 1), gen_i2s='resUnet', gen_s2i='resUnet')
    dI, dS = torch.tensor(I).cuda(), torch.tensor(Sg).cuda()
    for _ in range(3):
        gan.train_step(dI, dS)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        gan.train_step(dI, dS)
    torch.cuda.synchronize()
    print("S=%d b=1: %.1f ms/step wall" % (S, (time.perf_counter() - t0) / n * 1e3), flush=True)
    del gan
