"""On-device input pipeline (dataset.py:205-251): crop / flip / rot90 kernel against the numpy restatement for every
augmentation state, the segmentation retry loop, and sharded batches from .npy files."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _Args:
    DIMENSIONS, CHANNELS = 3, 1

    def __init__(self, S, G):
        self.SUBVOL_PATCH_SIZE, self.GLOBAL_BATCH_SIZE = (S, S, S), G


def test_crop_augment_matches_numpy_for_every_state(cuda):
    from oracle import np_ref
    from van_gan_b200.dataset import DatasetGen
    rng = np.random.default_rng(3)
    vol = rng.random((40, 37, 45)).astype(np.float32)
    ds = DatasetGen(_Args(16, 1), {}, {})
    dv = torch.from_numpy(vol).cuda()
    for size in ((16, 16, 16), (9, 12, 12), (8, 10, 14)):
        for flip_lr in (False, True):
            for flip_ud in (False, True):
                for k in (-1, 0, 1, 2, 3):
                    if k % 2 and size[1] != size[2]:
                        continue
                    origin = (5, 7, 11)
                    ref = np_ref.crop_augment(vol, origin, size, flip_lr, flip_ud, k)
                    got = ds.crop_augment(dv, origin, size, flip_lr, flip_ud, k).cpu().numpy()
                    assert np.array_equal(got, ref), (size, flip_lr, flip_ud, k)


def test_process_domains_and_retry(cuda):
    from oracle import np_ref
    from van_gan_b200.dataset import DatasetGen
    rng = np.random.default_rng(4)
    S = 16
    vol = (rng.random((48, 48, 48)) * 0.5).astype(np.float32)      # everywhere below SEG_THRESH ...
    vol[40:44, 40:44, 40:44] = 0.95                                # ... except one bright corner
    ds = DatasetGen(_Args(S, 1), {}, {}, seed=11)
    out = ds.process_seg_domain(vol[..., None])
    assert tuple(out.shape) == (S, S, S, 1)
    origin, flip_lr, flip_ud, k = ds.last_draw
    assert float(out.max()) >= ds.SEG_THRESH and ds.last_retries >= 1      # the accepted crop contains the bright corner
    assert k in (-1, 0)                                                     # the reference's radians // 90
    assert np.array_equal(out[..., 0].cpu().numpy(), np_ref.crop_augment(vol, origin, (S, S, S), flip_lr, flip_ud, k))
    img = ds.process_imaging_domain(vol[..., None])
    origin, flip_lr, flip_ud, k = ds.last_draw
    assert k == 0                                                           # preserve_depth_orientation=True
    assert np.array_equal(img[..., 0].cpu().numpy(), np_ref.crop_augment(vol, origin, (S, S, S), flip_lr, flip_ud, 0))
    # a volume that never reaches the threshold: the loop stops after 200 re-draws and returns the last crop (dataset.py:236-246)
    dark = DatasetGen(_Args(S, 1), {}, {}, seed=12)
    dark.process_seg_domain((vol * 0.1)[..., None])
    assert dark.last_retries == 200


def test_batches_from_npy_files(cuda, tmp_path):
    from van_gan_b200.dataset import DatasetGen
    from van_gan_b200.utils import process_imaging_otf
    rng = np.random.default_rng(5)
    S, G = 16, 4
    paths_i, paths_s = [], []
    for i in range(6):
        a = rng.random((32, 32, 32, 1)).astype(np.float32)
        b = (rng.random((32, 32, 32, 1)) > 0.7).astype(np.float32)
        np.save(tmp_path / ("i%d.npy" % i), a); np.save(tmp_path / ("s%d.npy" % i), b)
        paths_i.append(str(tmp_path / ("i%d.npy" % i))); paths_s.append(str(tmp_path / ("s%d.npy" % i)))
    ds = DatasetGen(_Args(S, G), {"training": paths_i}, {"training": paths_s}, otf_imaging=process_imaging_otf, seed=1)
    it = ds.batches('training')
    for _ in range(3):
        xi, xs = next(it)
        assert xi.is_cuda and tuple(xi.shape) == (G, S, S, S, 1) and tuple(xs.shape) == (G, S, S, S, 1)
        mn, mx = xi.amin(dim=(1, 2, 3, 4)), xi.amax(dim=(1, 2, 3, 4))
        assert torch.allclose(mn, -torch.ones_like(mn)) and torch.allclose(mx, torch.ones_like(mx))   # process_imaging_otf (main.py:169-177)
        assert float(xs.max()) >= 0.8
