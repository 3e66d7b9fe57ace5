import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nets as ON
from van_gan_b200 import engine as E
from van_gan_b200._lib import ACT_LEAKY, PAD_REFLECT, PAD_ZERO
from van_gan_b200.discriminator import get_discriminator
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b))
S = 32
rng = np.random.default_rng(3)
x = torch.tensor(rng.standard_normal((1, S, S, S, 1)), dtype=torch.float32).clamp(-1, 1)
initd = ON.init_params(ON.disc_param_shapes(), 3, 0.05)
nz, mk = ON.make_disc_rand(rng, 1, S)
d = get_discriminator((S, S, S, 1), filters=64, use_dropout=True, use_input_noise=True, use_layer_noise=True, name='d'); d.load(initd)
P = ON.to_torch(initd, requires_grad=False)
taps = {}
yo = ON.disc_forward(P, x, nz, mk, taps=taps)
print("oracle taps:", {k: tuple(v.shape) for k, v in taps.items()})
tape = E.Tape(enabled=False)
noise = [t.cuda() for t in nz]; masks = [m.cuda() for m in mk]
h0 = E.pad_noise(tape, E.Var(x.cuda()), noise=noise[0])
c0 = d.conv0(tape, h0)
ref_c0 = ON.conv3d(ON.reflect_pad(x) + nz[0], P["d0.conv.w"], P["d0.conv.b"], stride=2)
print("conv0", rel(c0.data.float(), ref_c0))
n0 = d.norm0(tape, c0, act=ACT_LEAKY, pad=(1, 1, PAD_REFLECT), noise=noise[1])
ref_n0 = ON.reflect_pad(torch.nn.functional.leaky_relu(ON.instance_norm(c0.data.float().cpu(), P["d0.in.gamma"], P["d0.in.beta"]), 0.2)) + nz[1]
print("norm0", rel(n0.data.float(), ref_n0))
c1 = d.conv1(tape, n0)
ref_c1 = ON.conv3d(n0.data.float().cpu(), P["d1.conv.w"].to(torch.bfloat16).float(), None, stride=2)
print("conv1 (tc s2)", rel(c1.data.float(), ref_c1), tuple(c1.shape))
out = d.forward(E.Tape(enabled=False), E.Var(x.cuda()), training=True, noise=noise, masks=masks)
print("full", rel(out.data, yo))
ref_nonoise = ON.reflect_pad(torch.nn.functional.leaky_relu(ON.instance_norm(c0.data.float().cpu(), P["d0.in.gamma"], P["d0.in.beta"]), 0.2))
print("norm0 vs ref without noise", rel(n0.data.float(), ref_nonoise))
diff = (n0.data.float().cpu() - ref_nonoise)
print("corr(diff, noise) =", float((diff * nz[1]).sum() / (diff.norm() * nz[1].norm())), "|diff|", float(diff.norm()), "|noise|", float(nz[1].norm()))
# shifted-noise hypotheses
flat_d, flat_n = diff.reshape(-1), nz[1].reshape(-1)
for sh in (0, 8, 64, -8):
    print("shift", sh, float((flat_d * torch.roll(flat_n, sh)).sum() / (flat_d.norm() * flat_n.norm())))
