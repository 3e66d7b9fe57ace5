"""Helpers the hot path uses from the reference's utils.py: min_max_norm (utils.py:10-24),
min_max_norm_tf (utils.py:27-48) and append_dict (utils.py:319-350)."""
import numpy as np
import torch

from . import engine as E
from ._lib import call


def min_max_norm(data):
    dmin, dmax = np.min(data), np.max(data)
    if (dmax - dmin) == 0:
        raise ValueError("Cannot perform min-max normalization when max and min are equal.")
    return (data - dmin) / (dmax - dmin)


def min_max_norm_tf(arr, axis=None):
    """Per-sample (axis=(1,2,3,4)) or whole-tensor (axis=None) min-max normalisation on the GPU."""
    arr = arr.contiguous()
    n = 1 if axis is None else arr.shape[0]
    v = arr.numel() // n
    mm = torch.empty(2 * n, dtype=torch.float32, device=E.DEV)
    enc = torch.empty(2 * n, dtype=torch.int32, device=E.DEV)
    call("vg_minmax", arr, n, v, mm, enc)
    out = torch.empty_like(arr)
    call("vg_minmax_normalize", arr, mm, out, n, v)
    return out


def append_dict(dict1, dict2, replace=False):
    for k, v in dict2.items():
        if replace or k not in dict1:
            dict1[k] = v if replace else [v]
        else:
            dict1[k].append(v)


def process_imaging_otf(tensor, axis=(1, 2, 3, 4), keepdims=True):
    """main.py:169-177: min/max normalisation of the imaging domain to [-1, 1], per sample (axis=(1,2,3,4)) or over the whole
    tensor (axis=None), on the device.  Accepted wherever the reference passes `process_imaging_domain=process_imaging_otf`."""
    t = torch.as_tensor(tensor, dtype=torch.float32, device=E.DEV).contiguous()
    out = min_max_norm_tf(t, axis=None if axis is None else (1, 2, 3, 4))
    res = torch.empty_like(out)
    call("vg_lincomb", res, res.numel(), 0, -1.0, out, 2.0, None, 0.0, None, 0.0)      # 2 * minmax(x) - 1
    return res


process_imaging_otf._vg_device = True
