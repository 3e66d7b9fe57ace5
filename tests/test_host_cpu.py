"""CPU: host-side logic that needs no GPU -- initializers, the numpy restatement of the input pipeline, the communicator's
run-time NCCL resolution (no collective is issued)."""
import ctypes
import math

import numpy as np


def test_default_init_follows_the_reference_initializers():
    from van_gan_b200 import engine as E
    shapes = {"stem.conv0.w": (3, 3, 3, 1, 16), "stem.conv0.b": (16,), "enc1.cb1.conv.w": (3, 3, 3, 16, 32), "enc1.cb1.in.gamma": (32,),
              "enc1.cb1.in.beta": (32,), "head.w": (1, 1, 1, 16, 1), "d1.in.gamma": (128,)}
    init = E.default_init(shapes, 7, glorot=("stem.conv0.w", "head.w", "d1.in.gamma"))
    lim = math.sqrt(6.0 / (27 * 1 + 27 * 16))                       # glorot_uniform: Conv3D built without kernel_initializer
    assert np.abs(init["stem.conv0.w"]).max() <= lim and np.abs(init["stem.conv0.w"]).max() > 0.8 * lim
    assert np.abs(init["head.w"]).max() <= math.sqrt(6.0 / 17)
    g = init["d1.in.gamma"]                                          # InstanceNormalization(gamma_initializer=None)
    assert np.abs(g).max() <= math.sqrt(3.0 / 128) and g.std() > 0
    w = init["enc1.cb1.conv.w"]                                      # he_normal: truncated at 2 sigma, std sqrt(2 / fan_in)
    sigma = math.sqrt(2.0 / (27 * 16))
    assert abs(w.std() / sigma - 1.0) < 0.05 and np.abs(w).max() <= 2.0 * sigma / 0.87962566103423978 + 1e-6
    assert np.all(init["enc1.cb1.in.gamma"] == 1) and np.all(init["enc1.cb1.in.beta"] == 0) and np.all(init["stem.conv0.b"] == 0)


def test_crop_augment_restatement_against_explicit_loops():
    """oracle.np_ref.crop_augment (dataset.py:205-230 under tf.image's 4-D reading of a volume) against index arithmetic."""
    from oracle import np_ref
    rng = np.random.default_rng(1)
    vol = rng.random((9, 8, 8)).astype(np.float32)
    n = 5
    for flip_lr in (False, True):
        for flip_ud in (False, True):
            for k in (-1, 0, 1, 2):
                out = np_ref.crop_augment(vol, (2, 1, 3), (4, n, n), flip_lr, flip_ud, k)
                for x in range(4):
                    for y in range(n):
                        for z in range(n):
                            yy, zz = {0: (y, z), 1: (z, n - 1 - y), 2: (n - 1 - y, n - 1 - z), 3: (n - 1 - z, y)}[k % 4]
                            if flip_ud:
                                yy = n - 1 - yy
                            if flip_lr:
                                zz = n - 1 - zz
                            assert out[x, y, z] == vol[2 + x, 1 + yy, 3 + zz]


def test_comm_resolves_nccl_at_run_time():
    import torch  # noqa: F401  (loads torch's bundled libnccl.so.2 into the process, as in any data-parallel run)
    from van_gan_b200 import _lib
    L = _lib.lib()
    assert L.vg_comm_available() == 1
    assert L.vg_comm_nccl_version() >= 21800
    buf = ctypes.create_string_buffer(128)
    assert L.vg_comm_unique_id(buf) == 0 and any(buf.raw)
    assert L.vg_comm_world(None) == 0 and L.vg_comm_rank(None) == -1 and L.vg_comm_destroy(None) == 0


def test_strategy_single_process_is_a_no_op():
    import torch
    from van_gan_b200.distribute import Strategy
    s = Strategy()
    assert s.num_replicas_in_sync == 1 and s.rank == 0
    t = torch.ones(4)
    assert s.all_reduce_async(t) is None and torch.equal(s.reduce("SUM", t), torch.ones(4))


def test_skeleton_ring_fixed_point_split_is_exact_and_order_independent():
    """The arithmetic of the soft-skeleton routing kernel's accumulator ring (csrc/skel.cu, skel_bwd_route_march_kernel), restated in
    numpy fp32: with 2^e above the largest magnitude, v = hi * 2^-sh + lo * 2^-(sh+25) exactly (sh = 25 - e), |hi| <= 2^25,
    |lo| <= 2^24, so 46 contributions fit an int32 and their integer sums do not depend on the order of the atomics."""
    rng = np.random.default_rng(41)
    f32 = np.float32
    for scale in (1e-20, 3e-7, 1.0, 777.0, 1e20):
        v = (rng.standard_normal(4096) * rng.choice([1e-6, 1e-3, 1.0], 4096) * scale).astype(f32)
        m = f32(np.abs(v).max())
        bits = np.array([m], dtype=f32).view(np.uint32)[0]
        e = int((bits >> 23) & 0xff) - 126                     # m = f * 2^e, f in [0.5, 1)
        assert m < 2.0 ** e and m >= 2.0 ** (e - 1)
        sh = max(-100, min(100, 25 - e))
        fx_hi, fx_hi_inv = f32(2.0 ** sh), f32(2.0 ** -sh)
        hs = np.rint(v * fx_hi).astype(f32)                    # exact: power-of-two scaling, |.| <= 2^25 is an integer-valued float
        assert np.abs(hs).max() <= 2 ** 25
        rem = ((v.astype(np.float64) - hs.astype(np.float64) * float(fx_hi_inv)).astype(f32) * fx_hi)   # the fmaf of the kernel
        assert np.array_equal(rem.astype(np.float64), (v.astype(np.float64) - hs.astype(np.float64) * 2.0 ** -sh) * 2.0 ** sh)
        assert np.abs(rem).max() <= 0.5
        lo = np.rint(rem * f32(2.0 ** 25)).astype(np.int64)
        hi = hs.astype(np.int64)
        assert np.abs(lo).max() <= 2 ** 24
        back = hi.astype(np.float64) * 2.0 ** -sh + lo.astype(np.float64) * 2.0 ** -(sh + 25)
        assert np.abs(back - v.astype(np.float64)).max() <= 2.0 ** (e - 51)          # half a unit of the low part
        # sums of 46 contributions per cell, in two different orders, in int32
        cells = rng.integers(0, 64, 46 * 64)
        idx = rng.integers(0, v.size, cells.size)
        acc_a, acc_b = np.zeros((2, 64), np.int64), np.zeros((2, 64), np.int64)
        for order, acc in ((np.arange(cells.size), acc_a), (rng.permutation(cells.size), acc_b)):
            np.add.at(acc[0], cells[order], hi[idx[order]])
            np.add.at(acc[1], cells[order], lo[idx[order]])
        assert np.array_equal(acc_a, acc_b) and np.abs(acc_a).max() < 2 ** 31
