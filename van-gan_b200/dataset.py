"""DatasetGen — the reference's input pipeline (dataset.py:10-251) with the per-sample work on the device.

Kept from the reference: the constructor arguments, `imaging_datagen` / `segmentation_datagen` (shuffled passes over lists of .npy
volumes, dataset.py:117-192), `process_imaging_domain` (random crop + flips, depth orientation preserved, :224-230),
`process_seg_domain` (random crop, re-drawn up to 200 times until the crop's maximum reaches SEG_THRESH = 0.8, then flips + rot90,
:232-251), `random_spatial_augmentation` (:205-222) and the `otf_imaging` hook applied to whole imaging batches (:58-60).
What changed: a volume is uploaded once and cropped / flipped / rotated by one kernel (`vg_crop_augment`), the retry test reads the
crop's maximum from the device (`vg_minmax`), and the batches that come out are CUDA tensors, already sharded per replica.
tf.data's prefetching / AUTOTUNE machinery is replaced by plain Python generators.  The random draws are numpy's (TensorFlow's
stateful RNG streams cannot be reproduced); what each draw MEANS follows the reference, including its quirks (see below).
"""
import math

import numpy as np
import torch

from . import engine as E
from ._lib import call


class DatasetGen:
    def __init__(self, args, imaging_domain_data, seg_domain_data, strategy=None, otf_imaging=None, semi_supervised_dir=None,
                 seed=0):
        if args.DIMENSIONS != 3:
            raise NotImplementedError("only DIMENSIONS=3 is built (main.py:80)")
        if semi_supervised_dir is not None:
            raise NotImplementedError("semi-supervised pairing (dataset.py:181-187) is out of scope")
        sp = args.SUBVOL_PATCH_SIZE
        self.imaging_patch_shape = (sp[0], sp[1], sp[2], args.CHANNELS)
        self.segmentation_patch_shape = (sp[0], sp[1], sp[2], 1)
        self.strategy = strategy
        self.imaging_paths = imaging_domain_data          # {'training': [...], 'validation': [...]} lists of .npy files
        self.segmentation_paths = seg_domain_data
        self.args = args
        self.otf_imaging = otf_imaging
        self.semi_supervised = False
        self.IMAGE_THRESH = 0.5
        self.SEG_THRESH = 0.8
        self.GLOBAL_BATCH_SIZE = args.GLOBAL_BATCH_SIZE
        self.rng = np.random.default_rng(seed)
        self.last_draw = None                              # (origin, flip_lr, flip_ud, rot_k) of the most recent sample (tests)
        self.last_retries = 0
        self._cache = {}                                   # path -> device volume (a volume is uploaded once)

    # ------------------------------------------------------------------ file generators (dataset.py:117-192)
    def _datagen(self, paths, typ):
        files = list(paths[typ])
        G = self.GLOBAL_BATCH_SIZE
        self.rng.shuffle(files)
        it = 0
        while True:
            if it >= math.floor(len(files) // G):
                it = 0
                self.rng.shuffle(files)
            for filename in files[it * G:(it + 1) * G]:
                yield self._volume(filename)
            it += 1

    def imaging_datagen(self, typ='training'):
        return self._datagen(self.imaging_paths, typ)

    def segmentation_datagen(self, typ='training'):
        return self._datagen(self.segmentation_paths, typ)

    def _volume(self, filename):
        if filename not in self._cache:
            a = np.load(filename).astype(np.float32)
            if a.ndim == 3:
                a = a[..., None]
            assert a.shape[-1] == 1, "single-channel volumes (CHANNELS = 1, main.py:79)"
            self._cache[filename] = torch.from_numpy(np.ascontiguousarray(a[..., 0])).to(E.DEV)
        return self._cache[filename]

    # ------------------------------------------------------------------ per-sample processing
    def _device_volume(self, image):
        if torch.is_tensor(image) and image.is_cuda:
            return image if image.dim() == 3 else image[..., 0].contiguous()
        a = np.asarray(image, dtype=np.float32)
        return torch.from_numpy(np.ascontiguousarray(a[..., 0] if a.ndim == 4 else a)).to(E.DEV)

    def _random_origin(self, vol, size):
        """tf.image.random_crop: a uniform offset per axis in [0, dim - size]."""
        return tuple(int(self.rng.integers(0, vol.shape[i] - size[i] + 1)) for i in range(3))

    def _draw_augmentation(self, preserve_depth_orientation):
        """random_spatial_augmentation (dataset.py:205-222).  Two coin flips; unless the depth orientation is preserved, rot90 with
        k = floor(angle_in_radians / 90) for an angle uniform in (-pi, pi) -- the reference divides RADIANS by 90, so k is -1 for a
        negative angle and 0 otherwise (kept)."""
        flip_lr = bool(self.rng.random() > 0.5)
        flip_ud = bool(self.rng.random() > 0.5)
        k = 0
        if not preserve_depth_orientation:
            angle = float(self.rng.uniform(-180.0, 180.0)) * (math.pi / 180.0)
            k = int(angle // 90)
        return flip_lr, flip_ud, k

    def crop_augment(self, vol, origin, size, flip_lr=False, flip_ud=False, rot_k=0):
        out = torch.empty((size[0], size[1], size[2]), dtype=torch.float32, device=E.DEV)
        call("vg_crop_augment", vol, vol.shape[0], vol.shape[1], vol.shape[2], out, size[0], size[1], size[2], origin[0], origin[1],
             origin[2], int(flip_lr), int(flip_ud), int(rot_k) % 4)
        return out

    def random_spatial_augmentation(self, image, max_rotation_angle=180, preserve_depth_orientation=False):
        """On an already cropped (device) volume."""
        vol = self._device_volume(image)
        flip_lr, flip_ud, k = self._draw_augmentation(preserve_depth_orientation)
        return self.crop_augment(vol, (0, 0, 0), vol.shape, flip_lr, flip_ud, k)[..., None]

    def process_imaging_domain(self, image):
        """dataset.py:224-230: random crop + flips (rotation skipped: preserve_depth_orientation=True)."""
        vol = self._device_volume(image)
        origin = self._random_origin(vol, self.imaging_patch_shape)
        flip_lr, flip_ud, k = self._draw_augmentation(True)
        self.last_draw = (origin, flip_lr, flip_ud, k)
        return self.crop_augment(vol, origin, self.imaging_patch_shape, flip_lr, flip_ud, k)[..., None]

    def _crop_max(self, crop):
        mm = torch.empty(2, dtype=torch.float32, device=E.DEV)
        enc = torch.empty(64, dtype=torch.int32, device=E.DEV)
        call("vg_minmax", crop, 1, crop.numel(), mm, enc)
        return float(mm[1].item())

    def process_seg_domain(self, image):
        """dataset.py:232-251: re-draw the crop (at most 200 times) while its maximum is below SEG_THRESH, then flips + rot90."""
        vol = self._device_volume(image)
        size = self.segmentation_patch_shape
        origin = self._random_origin(vol, size)
        crop = self.crop_augment(vol, origin, size)
        i = 0
        while i < 200 and self._crop_max(crop) < self.SEG_THRESH:
            origin = self._random_origin(vol, size)
            crop = self.crop_augment(vol, origin, size)
            i += 1
        self.last_retries = i
        flip_lr, flip_ud, k = self._draw_augmentation(False)
        self.last_draw = (origin, flip_lr, flip_ud, k)
        return self.crop_augment(vol, origin, size, flip_lr, flip_ud, k)[..., None]

    # ------------------------------------------------------------------ batches
    def _local(self):
        world = self.strategy.num_replicas_in_sync if self.strategy is not None else 1
        rank = self.strategy.rank if self.strategy is not None else 0
        assert self.GLOBAL_BATCH_SIZE % world == 0
        b = self.GLOBAL_BATCH_SIZE // world
        return rank * b, b

    def batches(self, typ='training'):
        """Yields (imaging batch, segmentation batch): this replica's shard of the global batch, CUDA fp32 [b, S, S, S, 1].  Every
        rank walks the same shuffled file order (same seed) and processes only its own slice of each global batch."""
        gi, gs = self.imaging_datagen(typ), self.segmentation_datagen(typ)
        first, b = self._local()
        G = self.GLOBAL_BATCH_SIZE
        while True:
            vi = [next(gi) for _ in range(G)]
            vs = [next(gs) for _ in range(G)]
            xi = torch.stack([self.process_imaging_domain(v) for v in vi[first:first + b]])
            xs = torch.stack([self.process_seg_domain(v) for v in vs[first:first + b]])
            if self.otf_imaging is not None:
                xi = self.otf_imaging(xi)
            yield xi, xs
