"""GPU tests of the callers either side of the hot path: the order-exact sliding-window stitcher (byte output must be EXACT),
GanMonitor's epoch callbacks (learning-rate schedule, discriminator-noise decay, plotter forward passes, per-window
process_imaging_domain hook), checkpoints with optimizer slots (save -> resume is bit-identical), run_mapping and epoch_sweep."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _TanhGen:
    def __call__(self, win, training=False):
        return torch.tanh(1.5 * torch.as_tensor(win, dtype=torch.float32, device="cuda") - 0.3)


def _np_gen(gen):
    return lambda t: gen(torch.tensor(np.ascontiguousarray(t))).cpu().numpy()


@pytest.mark.parametrize("shape,k,stride,complete", [((40, 40, 40), 16, (8, 8, 8), False), ((40, 40, 40), 16, (8, 8, 8), True),
                                                     ((37, 29, 16), 16, (7, 5, 1), True), ((48, 32, 32), 16, (16, 16, 16), False),
                                                     ((33, 47, 21), 16, (5, 9, 4), False)])
def test_stitch_is_bit_identical_to_the_numpy_loop(cuda, shape, k, stride, complete):
    """Same generator on both sides -> the uint8 (complete=False) and float32 (complete=True) results must be EQUAL, not close:
    every voxel adds its windows in the reference's enumeration order (custom_callback.py:142-192)."""
    from oracle import np_ref
    from van_gan_b200.custom_callback import GanMonitor
    rng = np.random.default_rng(11)
    vol = (rng.random(shape + (1,)) * 2 - 1).astype(np.float32)
    gen = _TanhGen()
    ref = np_ref.stitch_subvolumes(_np_gen(gen), vol, (1, k, k, k, 1), stride=stride, complete=complete, padFactor=0.25)
    for wb in (1, 3):
        mon = GanMonitor(window_batch=wb)
        got = mon.stitch_subvolumes(gen, vol, (1, k, k, k, 1), stride=stride, complete=complete, padFactor=0.25)
        assert got.dtype == ref.dtype and got.shape == ref.shape
        assert np.array_equal(got, ref), (wb, int((got != ref).sum()))
        assert mon.last_stats["unique"] <= mon.last_stats["windows"]


def test_stitch_runs_each_unique_window_once(cuda):
    from van_gan_b200.custom_callback import GanMonitor
    calls = []

    class Counting(_TanhGen):
        def __call__(self, win, training=False):
            calls.append(int(win.shape[0]))
            return super().__call__(win, training)
    vol = np.random.default_rng(1).random((48, 48, 32, 1)).astype(np.float32)
    mon = GanMonitor(window_batch=4)
    mon.use_graph = False
    mon.stitch_subvolumes(Counting(), vol, (1, 16, 16, 16, 1), stride=(16, 16, 16), complete=False)
    # (48-16)/16+1 = 3 (+1 repeated) per axis in H, W; 2 (+1) in D: 4*4*3 = 48 enumerated, 3*3*2 = 18 unique
    assert mon.last_stats == dict(windows=48, unique=18, local_windows=18)
    assert sum(calls) == 18


def test_stitch_process_img_hook(cuda):
    """custom_callback.py:171-172: the hook runs once per window; the device hook (utils.process_imaging_otf) and an equivalent
    host callable give the same result, which differs from the un-hooked one."""
    from oracle import np_ref
    from van_gan_b200.custom_callback import GanMonitor
    from van_gan_b200.utils import process_imaging_otf
    vol = (np.random.default_rng(2).random((32, 32, 16, 1)) * 3 + 1).astype(np.float32)
    gen = _TanhGen()

    def host_hook(arr, axis=None, keepdims=False):
        a = np.asarray(arr, np.float32)
        mx, mn = a.max(), a.min()
        return np.float32(2.0) * (a - mn) / (mx - mn) - np.float32(1.0)

    args = types.SimpleNamespace(INPUT_IMG_SIZE=(1, 16, 16, 16, 1), DIMENSIONS=3, output_dir="/tmp")
    a = GanMonitor(args, process_imaging_domain=process_imaging_otf, window_batch=3).stitch_subvolumes(
        gen, vol, (1, 16, 16, 16, 1), stride=(8, 8, 8), complete=False, process_img=True)
    b = GanMonitor(args, process_imaging_domain=host_hook, window_batch=3).stitch_subvolumes(
        gen, vol, (1, 16, 16, 16, 1), stride=(8, 8, 8), complete=False, process_img=True)
    c = GanMonitor(args, window_batch=3).stitch_subvolumes(gen, vol, (1, 16, 16, 16, 1), stride=(8, 8, 8), complete=False)
    assert np.abs(a.astype(int) - b.astype(int)).max() <= 1          # two float32 evaluation orders of 2*(x-mn)/(mx-mn)-1
    assert not np.array_equal(a, c)
    ref = np_ref.stitch_subvolumes(lambda t: _np_gen(gen)(host_hook(t[0])[None]), vol, (1, 16, 16, 16, 1), stride=(8, 8, 8), complete=False)
    assert np.array_equal(b, ref)


class _Args:
    def __init__(self, S, G, out):
        self.N_DEVICES, self.GLOBAL_BATCH_SIZE = 1, G
        self.INPUT_IMG_SIZE = (G, S, S, S, 1)
        self.CHANNELS, self.DIMENSIONS = 1, 3
        self.SUBVOL_PATCH_SIZE = (S, S, S)
        self.train_steps, self.BATCH_SIZE, self.output_dir = 4, G, out
        self.EPOCHS, self.INITIAL_LR, self.INITIATE_LR_DECAY, self.NO_NOISE = 10, 2e-4, 5, 10
        self.PERIOD_2D_CALLBACK, self.PERIOD_3D_CALLBACK = 2, 2


def _gan(tmp, S=32, G=1, seed=5):
    from van_gan_b200.vangan import VanGan
    return VanGan(_Args(S, G, str(tmp)), gen_i2s='resUnet', gen_s2i='resUnet', seed=seed)


def test_checkpoint_resume_is_bit_identical(cuda, tmp_path):
    """tf.train.Checkpoint holds the four models AND the four optimizers (vangan.py:238-245): after save -> load into a fresh
    trainer, the next step must give the same losses and the same weights as the uninterrupted run (Adam slots, iteration counts
    and the step counter that keys the noise streams all restored)."""
    from test_gpu_train_step import synth
    rng = np.random.default_rng(3)
    batches = [synth(rng, 1, 32) for _ in range(3)]
    a = _gan(tmp_path)
    a.use_graph = False
    for I, Sg in batches[:2]:
        a.train_step(I.cuda(), Sg.cuda())
    path = a.save_checkpoint(epoch=6)
    assert path.endswith(os.path.join("checkpoints", "checkpoint_e7.npz")) and os.path.exists(path)
    ra = a.train_step(batches[2][0].cuda(), batches[2][1].cuda())
    b = _gan(tmp_path)                             # same seed (it keys the noise streams) ...
    b.use_graph = False
    for n in b.networks.values():                  # ... but different weights and slots: everything must come from the checkpoint
        n.w.mul_(0.3); n.m.fill_(1.0); n.v.fill_(2.0)
    assert b.load_checkpoint(epoch=7) is True
    assert b.step == 2 and all(n.step_count == 2 for n in b.networks.values())
    rb = b.train_step(batches[2][0].cuda(), batches[2][1].cuda())
    for k in ra:
        assert abs(ra[k] - rb[k]) <= 1e-6 * abs(ra[k]), (k, ra[k], rb[k])
    for k in a.networks:
        d = float((a.networks[k].w - b.networks[k].w).abs().max())
        assert d <= 1e-6, (k, d)                   # fp32 atomics order in the weight-gradient kernels is the only difference (measured 2.4e-7)
    assert b.load_checkpoint(epoch=123) is False   # prints the reference's "Checkpoint not found" line
    z = np.load(path)
    assert "gen_IS/stem.conv0.w" in z.files and "gen_I_optimizer/m/stem.conv0.w" in z.files and "disc_S_optimizer/iter" in z.files


def test_epoch_callbacks_lr_schedule_and_noise_decay(cuda, tmp_path):
    from van_gan_b200 import engine as E
    from van_gan_b200.custom_callback import GanMonitor
    from test_gpu_train_step import synth
    gan = _gan(tmp_path)
    args = _Args(32, 1, str(tmp_path))
    mon = GanMonitor(args)
    I, Sg = synth(np.random.default_rng(4), 1, 32)
    mon.on_epoch_start(gan, 0, args)
    assert gan.disc_I.noise_std == pytest.approx(0.1) and gan.gen_I_optimizer.current_lr() == pytest.approx(2e-4)
    for _ in range(4):
        gan.train_step(I.cuda(), Sg.cuda())
    assert gan._graph is not None
    mon.on_epoch_start(gan, 4, args)               # noise 0.1 * (1 - 4/10)
    assert gan.disc_I.noise_std == pytest.approx(0.06) and gan.disc_S.noise_std == pytest.approx(0.06)
    gan.train_step(I.cuda(), Sg.cuda())
    assert gan._graph is not None and gan._graph["noise"] == (pytest.approx(0.06), pytest.approx(0.06))   # re-captured
    mon.on_epoch_start(gan, 5, args)               # == INITIATE_LR_DECAY: PolynomialDecay over (10-5)*4 = 20 steps of the iteration counter
    sched = gan.gen_I_optimizer.lr
    assert isinstance(sched, E.PolynomialDecay) and sched.decay_steps == 20
    it = gan.gen_I_optimizer.iterations
    assert it == 5 and gan.gen_I_optimizer.current_lr() == pytest.approx(2e-4 * (1 - 5 / 20))
    w0 = gan.gen_IS.w.clone()
    gan.train_step(I.cuda(), Sg.cuda())
    assert gan.gen_I_optimizer.current_lr() == pytest.approx(2e-4 * (1 - 6 / 20))
    assert float((gan.gen_IS.w - w0).abs().max()) > 0
    for o in gan.optimizers.values():              # lr 0 from here on: weights must stop moving
        o.lr = 0.0
    w1 = {k: n.w.clone() for k, n in gan.networks.items()}
    gan.train_step(I.cuda(), Sg.cuda())
    assert all(torch.equal(w1[k], n.w) for k, n in gan.networks.items())
    mon.updateDiscriminatorNoise(gan.disc_I, gan.layer_noise, 20, args)
    assert gan.disc_I.noise_std == 0.0


def test_image_plotter_forward_passes(cuda, tmp_path, monkeypatch):
    """custom_callback.py:265-267: prediction = genX(sample), cycled = genY(prediction), identity = genY(sample) on a random crop."""
    from van_gan_b200.custom_callback import GanMonitor
    monkeypatch.chdir(tmp_path)
    gan = _gan(tmp_path)
    args = _Args(32, 1, str(tmp_path))
    vol = (np.random.default_rng(6).random((40, 36, 34, 1)) * 2 - 1).astype(np.float32)
    ds = types.SimpleNamespace(imaging_val_full_vol_data=[(vol, 0)], segmentation_val_full_vol_data=[(vol, 0)])
    mon = GanMonitor(args, dataset=ds, imaging_val_data=["/x/sampleA.npy"], segmentation_val_data=["/x/sampleB.npy"])
    pa, pb = mon.on_epoch_end(gan, 1)
    assert pa["name"] == "sampleA" and pb["name"] == "sampleB"
    for k in ("sample", "prediction", "cycled", "identity"):
        assert pa[k].shape == (32, 32, 32, 1) and np.isfinite(pa[k]).all()
    x = torch.tensor(pa["sample"][None]).cuda()
    pred = gan.gen_IS(x, training=False)
    assert np.array_equal(pa["prediction"], pred[0].cpu().numpy())
    assert np.array_equal(pa["cycled"], gan.gen_SI(pred, training=False)[0].cpu().numpy())
    assert np.array_equal(pa["identity"], gan.gen_SI(x, training=False)[0].cpu().numpy())


def test_run_mapping_and_epoch_sweep(cuda, tmp_path):
    """custom_callback.py:466-509 and post_training.py:4-39 on .npy volumes: checkpoints are restored per epoch, every test file is
    mapped into <output_dir>/Epoch_Sampling/e{i}/e{i}_VG_<name>.npy, missing epochs print the error line and keep going."""
    from van_gan_b200.custom_callback import GanMonitor
    from van_gan_b200.post_training import epoch_sweep
    gan = _gan(tmp_path)
    args = _Args(32, 1, str(tmp_path))
    test_dir = tmp_path / "test"
    test_dir.mkdir()
    rng = np.random.default_rng(8)
    for name in ("volA", "volB"):
        np.save(test_dir / (name + ".npy"), (rng.random((40, 40, 40, 1)) * 2 - 1).astype(np.float32))
    gan.save_checkpoint(epoch=1)                   # -> checkpoint_e2
    w_e2 = gan.gen_IS.w.clone()
    gan.gen_IS.w.mul_(0.5); gan.gen_IS.repack()    # perturb, then save another epoch
    gan.save_checkpoint(epoch=3)                   # -> checkpoint_e4
    mon = GanMonitor(args, window_batch=2)
    out = epoch_sweep(args, gan, mon, test_path=str(test_dir), start=2, end=4, step=2, segmentation=True)
    assert sorted(out) == [2, 4] and len(out[2]) == 2
    for i in (2, 4):
        for name in ("volA", "volB"):
            f = tmp_path / "Epoch_Sampling" / ("e%d" % i) / ("e%d_VG_%s.npy" % (i, name))
            assert f.exists()
            arr = np.load(f)
            assert arr.shape == (40, 40, 40, 1) and arr.dtype == np.float32
            assert abs(float(arr.min())) < 1e-4 and abs(float(arr.max()) - 255.0) < 1e-3
    assert not np.array_equal(out[2][0], out[4][0])      # the two epochs really used different weights
    assert torch.equal(gan.gen_IS.w, w_e2 * 0.5)          # the last restore was epoch 4
    # run_mapping with the imaging-domain direction applies the hook per window (process_img=True, custom_callback.py:507-509)
    from van_gan_b200.utils import process_imaging_otf
    mon2 = GanMonitor(args, process_imaging_domain=process_imaging_otf, window_batch=2)
    res = mon2.run_mapping(gan, [str(test_dir / "volA.npy")], args.INPUT_IMG_SIZE, segmentation=False, stride=(16, 16, 16), padFactor=0.1,
                           filetext="m_", filepath=str(tmp_path))
    assert res[0].shape == (40, 40, 40, 1) and (tmp_path / "m_volA.npy").exists()
