import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_train_step import synth, rel_l2
from oracle import nets as ON
from van_gan_b200 import engine as E
from van_gan_b200.resunet_model import ResUNet
for S in (32, 64):
    rng = np.random.default_rng(3)
    real_I, real_S = synth(rng, 1, S)
    init = ON.init_params(ON.resunet_param_shapes(), 1, 0.05)
    P = ON.to_torch(init, requires_grad=False)
    net = ResUNet((S, S, S, 1), upsample_mode='simple')
    net.load(init)
    tk = {}
    out = net.forward(E.Tape(enabled=False), E.Var(real_I.cuda()), taps=tk)
    for emu in (True, False):
        ON.Emu.on = emu
        to = {}
        yo = ON.resunet_forward(P, real_I, taps=to)
        ON.Emu.on = False
        print("S", S, "emu", emu, " ".join("%s %.4f" % (k, rel_l2(tk[k].data.float().cpu(), to[k])) for k in to), "out %.4f" % rel_l2(out.data.cpu(), yo))
