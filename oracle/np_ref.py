"""Oracle, second opinion: numpy/scipy restatement of the stencil and stitching arithmetic.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Independent of torch so the two restatements
can be checked against each other bit for bit (min/max/sub/mul/relu are exactly reproducible).

Follows clDice_func.py:8-80 (soft_erode/dilate/open/skel), loss_functions.py:86-117 (ssim_loss_3d),
utils.py:10-24 (min_max_norm) and custom_callback.py:47-223 (stitch_subvolumes, 3-D branch).
Arrays are single volumes (D,H,W) unless stated.
"""
import numpy as np
from scipy import ndimage


def soft_erode(v):
    # 'same' pooling ignores out-of-volume voxels -> pad with +inf for a minimum filter
    kw = dict(mode="constant", cval=np.inf)
    p1 = ndimage.minimum_filter(v, size=(3, 3, 1), **kw)
    p2 = ndimage.minimum_filter(v, size=(3, 1, 3), **kw)
    p3 = ndimage.minimum_filter(v, size=(1, 3, 3), **kw)
    return np.minimum(np.minimum(p1, p2), p3)


def soft_dilate(v):
    return ndimage.maximum_filter(v, size=(3, 3, 3), mode="constant", cval=-np.inf)


def soft_skel(v, iters):
    relu = lambda t: np.maximum(t, 0).astype(v.dtype)
    img = v
    skel = relu(img - soft_dilate(soft_erode(img)))
    for _ in range(iters):
        img = soft_erode(img)
        delta = relu(img - soft_dilate(soft_erode(img)))
        skel = skel + relu(delta - skel * delta)
    return skel


def ssim_map(t, p, sigma=1.5, k1=0.01, k2=0.03):
    grid = np.arange(-1, 2, dtype=np.float64)
    g = np.exp(-0.5 * (grid / sigma) ** 2) / (sigma * np.sqrt(2 * np.pi))
    g = g / g.sum()
    w = np.einsum("i,j,k->ijk", g, g, g)
    blur = lambda a: ndimage.correlate(a.astype(np.float64), w, mode="constant", cval=0.0)
    mt, mp = blur(t), blur(p)
    stt, spp, stp = blur(t * t) - mt * mt, blur(p * p) - mp * mp, blur(t * p) - mt * mp
    c1, c2 = k1 ** 2, k2 ** 2
    return (2 * mt * mp + c1) * (2 * stp + c2) / ((mt * mt + mp * mp + c1) * (stt + spp + c2))


def min_max_norm(data):
    dmin, dmax = np.min(data), np.max(data)
    if (dmax - dmin) == 0:
        raise ValueError("Cannot perform min-max normalization when max and min are equal.")
    return (data - dmin) / (dmax - dmin)


def window_starts(n, k, s):
    """custom_callback.py:127-162 — dim_out+1 iterations, start clamped to n-k (so the last window
    is flush with the edge and is DUPLICATED when (n-k) is a multiple of the stride)."""
    dim_out = int(np.floor((n - k) / s + 1))
    out, start = [], 0
    for _ in range(dim_out + 1):
        if start > n - k:
            start = n - k
        out.append(start)
        start += s
    return out


def stitch_subvolumes(gen, img, subvol_size, stride=(25, 25, 128), complete=False, padFactor=0.25,
                      border_removal=True):
    """3-D branch of GanMonitor.stitch_subvolumes.  img: (H,W,D,C) float32; subvol_size: (N,kH,kW,kD,[C]);
    gen: callable (1,kH,kW,kD,C) -> (1,kH,kW,kD,C).  Returns the float array BEFORE the TIFF write
    (`255*min_max_norm(pred)`, cast to uint8 when complete=False, custom_callback.py:202-205)."""
    img = np.asarray(img, dtype=np.float32)
    oshape = img.shape
    if complete:
        xs, ys = int(padFactor * img.shape[0]), int(padFactor * img.shape[1])
        if stride[2] == 1:
            zs = 0
            img = np.pad(img, ((xs, xs), (ys, ys), (0, 0), (0, 0)), "symmetric")
        else:
            zs = int(padFactor * img.shape[2])
            img = np.pad(img, ((xs, xs), (ys, ys), (zs, zs), (0, 0)), "symmetric")
    H, W, D, C = img.shape
    kH, kW, kD = subvol_size[1], subvol_size[2], subvol_size[3]
    if not complete or not border_removal:
        pH = pW = pD = 0
    else:
        pH, pW, pD = int(0.1 * kH), int(0.1 * kW), int(0.1 * kD)
        if kD == D:
            pD = 0
    cnt = np.zeros((H, W, D, C), np.float32)
    pred = np.zeros(img.shape, np.float32)
    for r in window_starts(H, kH, stride[0]):
        for c in window_starts(W, kW, stride[1]):
            for d in window_starts(D, kD, stride[2]):
                sl = (slice(r + pH, r + kH - pH), slice(c + pW, c + kW - pW), slice(d + pD, d + kD - pD))
                cnt[sl] += 1.0
                out = np.asarray(gen(img[None, r:r + kH, c:c + kW, d:d + kD]))[0]
                pred[sl] += out[pH:kH - pH, pW:kW - pW, pD:kD - pD]
    with np.errstate(invalid="ignore", divide="ignore"):
        pred = np.true_divide(pred, cnt)
    if complete:
        if stride[2] == 1:
            pred = pred[xs:oshape[0] + xs, ys:oshape[1] + ys]
        else:
            pred = pred[xs:oshape[0] + xs, ys:oshape[1] + ys, zs:oshape[2] + zs]
    pred = 255 * min_max_norm(pred)
    if not complete:
        pred = pred.astype("uint8")
    return pred


def crop_augment(vol, origin, size, flip_lr=False, flip_ud=False, rot_k=0):
    """dataset.py:205-230 on a (H, W, D) volume: tf.image.random_crop at `origin`, then random_spatial_augmentation.  tf.image reads
    the 4-D volume tensor (H, W, D, 1) as [batch, height, width, channels]: flip_left_right reverses axis 2 (D), flip_up_down axis 1
    (W), rot90(k) turns the (W, D) plane k quarter turns counter-clockwise."""
    x0, y0, z0 = origin
    a = vol[x0:x0 + size[0], y0:y0 + size[1], z0:z0 + size[2]]
    if flip_lr:
        a = a[:, :, ::-1]
    if flip_ud:
        a = a[:, ::-1, :]
    return np.ascontiguousarray(np.rot90(a, rot_k, axes=(1, 2)))
