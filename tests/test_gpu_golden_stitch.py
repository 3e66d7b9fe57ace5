"""GPU: committed golden fixtures (tests/golden, produced by the CPU oracle) and the sliding-window stitcher."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().flatten(), torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_golden_losses(cuda):
    from van_gan_b200 import clDice_func as K, engine as E, loss_functions as LF
    z = np.load(os.path.join(GOLD, "losses_20.npz"))
    x01 = torch.tensor(z["x01"]).cuda()
    assert np.array_equal(K.soft_skel(x01, 5).cpu().numpy(), z["skel5"])          # bit-exact
    assert np.array_equal(K.soft_erode(x01).cpu().numpy(), z["erode"])

    class Cfg:
        global_batch_size, n_devices, lambda_cycle, lambda_reconstruction, lambda_topology = 2, 2, 10.0, 5.0, 5.0
    cfg = Cfg()
    fns = {"bce": lambda r, c: LF.cycle_loss(cfg, r, c, typ="bce"), "mse": lambda r, c: LF.cycle_loss(cfg, r, c, typ="mse"),
           "ssim": lambda r, c: LF.cycle_reconstruction(cfg, r, c), "seg": lambda r, c: LF.cycle_seg_loss(cfg, r, c, iters=5)}
    for k, fn in fns.items():
        cfg.loss_ctx = LF.LossContext()
        rv, cv = E.Var(torch.tensor(z["real"]).cuda()), E.Var(torch.tensor(z["cycled"]).cuda())
        s = fn(rv, cv)
        assert abs(float(s) - float(z["val_" + k])) <= 1e-5 * abs(float(z["val_" + k])), k    # fp32 ops: 1e-5
        g = sum(gg for v, gg in s.seeds() if v is cv).cpu()
        assert rel_l2(g, z["grad_" + k]) < 2e-5, k


def test_golden_networks(cuda):
    from oracle import nets as ON          # only to rebuild the seeded weights the fixture was made with
    from van_gan_b200 import engine as E
    from van_gan_b200.discriminator import get_discriminator
    from van_gan_b200.resunet_model import ResUNet
    z = np.load(os.path.join(GOLD, "nets_32.npz"))
    g = ResUNet((32, 32, 32, 1), upsample_mode='simple')
    g.load(ON.init_params(ON.resunet_param_shapes(), 7, 0.05))
    y = g(torch.tensor(z["x"]))
    assert rel_l2(y.cpu(), z["gen_out"]) < 8e-2      # bf16 storage, 58 rounding stages, 2^3-voxel norms at 32^3
    d = get_discriminator((32, 32, 32, 1), filters=64, use_dropout=True, use_input_noise=True, use_layer_noise=True, name='d')
    d.load(ON.init_params(ON.disc_param_shapes(), 8, 0.05))
    yd = d.forward(E.Tape(enabled=False), E.Var(torch.tensor(z["x"]).cuda()), training=True,
                   noise=[torch.tensor(z["noise%d" % i]).cuda() for i in range(5)],
                   masks=[torch.tensor(z["mask%d" % i]).cuda() for i in range(3)]).data
    assert rel_l2(yd.cpu(), z["disc_out"]) < 2e-2
    assert rel_l2(d(torch.tensor(z["x"])).cpu(), z["disc_out_inference"]) < 2e-2


class _TanhGen:
    """stand-in generator with exactly reproducible arithmetic, so the stitching itself can be compared tightly"""

    def __call__(self, win, training=False):
        return torch.tanh(1.5 * win - 0.3)


def test_stitch_against_golden_and_numpy(cuda):
    from oracle import np_ref
    from van_gan_b200.custom_callback import GanMonitor, window_starts
    z = np.load(os.path.join(GOLD, "stitch_40.npz"))
    mon = GanMonitor(window_batch=3)
    a = mon.stitch_subvolumes(_TanhGen(), z["vol"], (1, 16, 16, 16, 1), stride=(8, 8, 8), complete=True, padFactor=0.25)
    assert a.shape == z["complete"].shape and a.dtype == np.float32
    assert np.allclose(a, z["complete"], rtol=1e-5, atol=2e-3)          # values span 0..255
    b = mon.stitch_subvolumes(_TanhGen(), z["vol"], (1, 16, 16, 16, 1), stride=(8, 8, 8), complete=False)
    # the fixture was made with numpy's tanh on the host; CUDA's tanh differs in the last bit, so +-1 here -- the EXACT comparison
    # (same generator on both sides) is tests/test_gpu_monitor_ckpt.py::test_stitch_is_bit_identical_to_the_numpy_loop
    assert b.dtype == np.uint8 and np.abs(b.astype(int) - z["plain"].astype(int)).max() <= 1
    assert window_starts(100, 64, 25) == np_ref.window_starts(100, 64, 25)
    # ragged case: stride does not divide, last window clamped; depth equal to the window (pD = 0 branch)
    rng = np.random.default_rng(5)
    vol = rng.random((37, 29, 16, 1)).astype(np.float32)
    ref = np_ref.stitch_subvolumes(lambda t: np.tanh(1.5 * t - 0.3), vol, (1, 16, 16, 16, 1), stride=(7, 5, 1), complete=True,
                                   padFactor=0.25)
    got = mon.stitch_subvolumes(_TanhGen(), vol, (1, 16, 16, 16, 1), stride=(7, 5, 1), complete=True, padFactor=0.25)
    assert np.allclose(got, ref, rtol=1e-5, atol=2e-3)


def test_stitch_with_resunet_generator(cuda):
    """the real generator on the CUDA path inside the stitcher vs the numpy stitcher driven by the same generator"""
    from oracle import np_ref
    from van_gan_b200.custom_callback import GanMonitor
    from van_gan_b200.resunet_model import ResUNet
    rng = np.random.default_rng(6)
    vol = (rng.random((48, 48, 32, 1)) * 2 - 1).astype(np.float32)
    g = ResUNet((32, 32, 32, 1), upsample_mode='simple', seed=3)
    mon = GanMonitor(window_batch=4)
    got = mon.stitch_subvolumes(g, vol, (1, 32, 32, 32, 1), stride=(16, 16, 16), complete=True, padFactor=0.25)
    ref = np_ref.stitch_subvolumes(lambda t: g(torch.tensor(t)).cpu().numpy(), vol, (1, 32, 32, 32, 1), stride=(16, 16, 16),
                                   complete=True, padFactor=0.25)
    assert np.allclose(got, ref, rtol=1e-4, atol=5e-2)
    assert mon.last_stats["windows"] > mon.last_stats["unique"] > 0
