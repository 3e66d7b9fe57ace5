"""GanMonitor — the reference's sliding-window predictor (custom_callback.py:47-223, 466-509) with the
window loop, accumulation, overlap-count division and min-max rescale done on the device.

Kept from the reference: `stitch_subvolumes(gen, img, subvol_size, epoch, stride, name, output_path,
complete, padFactor, border_removal, process_img)` and `run_mapping(model, test_set, sub_img_size,
segmentation, stride, padFactor, filetext, filepath)`, the exact window enumeration (dim_out+1
iterations per axis with end clamping, so the last window is duplicated when (dim-k) is a multiple
of the stride), the 10 % border crop, UNIFORM overlap-count blending and the final
`255 * min_max_norm`.  Changed: windows are run through the generator in batches, volumes stay in
HBM, windows are sharded round-robin over the data-parallel ranks (one all-reduce of the partial
sums at the end).  The TIFF writers (skimage) are out of scope: results are returned, and saved as
.npy when an output path is given.
"""
import os

import numpy as np
import torch

from . import engine as E
from ._lib import call
from .distribute import Strategy


def window_starts(n, k, s):
    """custom_callback.py:127-162: dim_out+1 iterations, start clamped to n-k."""
    dim_out = int(np.floor((n - k) / s + 1))
    out, start = [], 0
    for _ in range(dim_out + 1):
        if start > n - k:
            start = n - k
        out.append(start)
        start += s
    return out


class GanMonitor:
    def __init__(self, args=None, dataset=None, Alabel="I", Blabel="S", model_path=None, imaging_val_data=None,
                 segmentation_val_data=None, process_imaging_domain=None, period=5, strategy=None, window_batch=4):
        self.dims = getattr(args, "DIMENSIONS", 3) if args is not None else 3
        if self.dims != 3:
            raise NotImplementedError("only DIMENSIONS=3 is built (main.py:80)")
        self.model_path = model_path
        self.process_imaging_domain = process_imaging_domain
        self.period = period
        self.strategy = strategy if strategy is not None else Strategy()
        self.window_batch = window_batch
        self.last_stats = None

    def stitch_subvolumes(self, gen, img, subvol_size, epoch=-1, stride=(25, 25, 128), name=None, output_path=None,
                          complete=False, padFactor=0.25, border_removal=True, process_img=False):
        """img: (H,W,D,1) float array (host).  gen: a generator model of this package (ResUNetModel).
        Returns the stitched prediction (float32, or uint8 when complete=False) exactly as the reference
        computes it before its TIFF write."""
        if process_img and self.process_imaging_domain is not None:
            raise NotImplementedError("per-window process_imaging_domain hook (host callback) is not built")
        img = np.asarray(img, dtype=np.float32)
        oshape = img.shape
        xs = ys = zs = 0
        if complete:                                                    # custom_callback.py:82-104
            xs, ys = int(padFactor * img.shape[0]), int(padFactor * img.shape[1])
            if stride[2] == 1:
                img = np.pad(img, ((xs, xs), (ys, ys), (0, 0), (0, 0)), "symmetric")
            else:
                zs = int(padFactor * img.shape[2])
                img = np.pad(img, ((xs, xs), (ys, ys), (zs, zs), (0, 0)), "symmetric")
        H, W, D, C = img.shape
        assert C == 1
        kH, kW, kD = subvol_size[1], subvol_size[2], subvol_size[3]
        if not complete or not border_removal:
            pH = pW = pD = 0
        else:
            pH, pW, pD = int(0.1 * kH), int(0.1 * kW), int(0.1 * kD)
            if kD == D:
                pD = 0
        starts = [(r, c, d) for r in window_starts(H, kH, stride[0]) for c in window_starts(W, kW, stride[1])
                  for d in window_starts(D, kD, stride[2])]
        rank, world = self.strategy.rank, self.strategy.num_replicas_in_sync
        mine = starts[rank::world]                                     # windows are independent: shard round-robin
        vol = torch.from_numpy(np.ascontiguousarray(img[..., 0])).to(E.DEV)
        pred = torch.zeros((H, W, D), dtype=torch.float32, device=E.DEV)
        cnt = torch.zeros((H, W, D), dtype=torch.float32, device=E.DEV)
        B = self.window_batch
        for i in range(0, len(mine), B):
            chunk = mine[i:i + B]
            st = torch.tensor(chunk, dtype=torch.int32, device=E.DEV).reshape(-1)
            win = torch.empty((len(chunk), kH, kW, kD, 1), dtype=torch.float32, device=E.DEV)
            call("vg_stitch_gather", vol, H, W, D, win, st, len(chunk), kH, kW, kD)
            out = gen(win, training=False)                             # batched generator forward on the CUDA path
            call("vg_stitch_accumulate", pred, cnt, H, W, D, out.contiguous(), st, len(chunk), kH, kW, kD, pH, pW, pD)
        if world > 1:
            self.strategy.reduce("SUM", pred)
            self.strategy.reduce("SUM", cnt)
        oH, oW, oD = (oshape[0], oshape[1], oshape[2]) if complete else (H, W, D)
        if complete and stride[2] == 1:
            oD = D
        out = torch.empty((oH, oW, oD), dtype=torch.float32, device=E.DEV)
        mm = torch.empty(2, dtype=torch.float32, device=E.DEV)
        enc = torch.empty(2, dtype=torch.int32, device=E.DEV)
        call("vg_stitch_finalize", pred, cnt, H, W, D, xs, ys, zs, oH, oW, oD, out, mm, enc)
        call("vg_stitch_scale", out, out.numel(), mm)
        self.last_stats = dict(windows=len(starts), unique=len(set(starts)), local_windows=len(mine))
        if not complete:
            out = out.to(torch.uint8)        # custom_callback.py:204-205 (astype('uint8')): cast on the device, 4x less D2H
        res = out.cpu().numpy()[..., None]
        if output_path is not None and name is not None and rank == 0:
            np.save(os.path.join(output_path, "{name}.npy".format(name=name)), res)
        return res

    def run_mapping(self, model, test_set, sub_img_size=(64, 64, 512, 1), segmentation=True, stride=(25, 25, 1),
                    padFactor=0.25, filetext=None, filepath=''):
        """custom_callback.py:466-509: every file of test_set (.npy volumes) through gen_IS (segmentation) or gen_SI."""
        results = []
        for imgdir in range(len(test_set)):
            img = np.load(test_set[imgdir])
            if img.ndim == 3:
                img = img[..., None]
            filename = os.path.splitext(os.path.basename(test_set[imgdir]))[0]
            gen = model.gen_IS if segmentation else model.gen_SI
            if not segmentation and self.process_imaging_domain is not None:
                raise NotImplementedError("process_img=True path (per-window host callback) is not built")
            results.append(self.stitch_subvolumes(gen, img, sub_img_size, name=(filetext or "") + filename,
                                                  output_path=filepath or None, complete=True, stride=stride,
                                                  padFactor=padFactor))
        return results
