"""Oracle: clDice soft-skeleton, cycle / SSIM / LSGAN losses (torch CPU, autograd).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  NDHWC tensors, fp32 or fp64.

Follows
  clDice_func.py:8-26 soft_erode   :29-42 soft_dilate   :45-57 soft_open   :60-80 soft_skel
  clDice_func.py:83-102 soft_clDice_loss   :105-119 soft_dice   :122-149 soft_dice_cldice_loss
  loss_functions.py:7-22 reduce_mean   :56-68 MSE   :86-117 ssim_loss_3d   :163-190 cycle_loss
  loss_functions.py:193-208 cycle_reconstruction   :211-226 cycle_seg_loss
  loss_functions.py:255-286 generator_loss_fn      :289-322 discriminator_loss_fn
  utils.py:27-48 min_max_norm_tf
`cfg` carries the attributes the reference reads from the VanGan instance:
global_batch_size, n_devices, lambda_cycle, lambda_reconstruction, lambda_topology.
"""
import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F


def make_cfg(global_batch_size, n_devices=1, lambda_cycle=10.0, lambda_reconstruction=5.0, lambda_topology=5.0):
    return SimpleNamespace(global_batch_size=global_batch_size, n_devices=n_devices, lambda_cycle=lambda_cycle,
                           lambda_reconstruction=lambda_reconstruction, lambda_topology=lambda_topology)


def _c(x):   # NDHWC -> NCDHW
    return x.permute(0, 4, 1, 2, 3)


def _l(x):
    return x.permute(0, 2, 3, 4, 1)


def _maxpool_same(x, k):
    """KL.MaxPool3D(pool_size=k, strides=1, padding='same'): out-of-volume voxels are ignored."""
    pad = tuple(s // 2 for s in k)
    return _l(F.max_pool3d(_c(x), kernel_size=k, stride=1, padding=pad))


def soft_erode(img):
    p1 = -_maxpool_same(-img, (3, 3, 1))
    p2 = -_maxpool_same(-img, (3, 1, 3))
    p3 = -_maxpool_same(-img, (1, 3, 3))
    return torch.minimum(torch.minimum(p1, p2), p3)


def soft_dilate(img):
    return _maxpool_same(img, (3, 3, 3))


def soft_open(img):
    return soft_dilate(soft_erode(img))


def soft_skel(img, iters):
    """Literal transcription of clDice_func.py:60-80 (2*iters+1 erodes)."""
    img1 = soft_open(img)
    skel = torch.relu(img - img1)
    for _ in range(iters):
        img = soft_erode(img)
        img1 = soft_open(img)
        delta = torch.relu(img - img1)
        intersect = skel * delta
        skel = skel + torch.relu(delta - intersect)
    return skel


def soft_skel_dedup(img, iters):
    """Same values with iters+1 erodes: the erode inside soft_open of iteration j is the image of
    iteration j+1 (bitwise identical; checked in tests/test_oracle_losses.py)."""
    e = [img]
    for _ in range(iters + 1):
        e.append(soft_erode(e[-1]))
    skel = torch.relu(e[0] - soft_dilate(e[1]))
    for j in range(1, iters + 1):
        delta = torch.relu(e[j] - soft_dilate(e[j + 1]))
        skel = skel + torch.relu(delta - skel * delta)
    return skel


def soft_clDice_loss(y_true, y_pred, iter_=50):
    smooth = 1.0
    skel_pred = soft_skel(y_pred, iter_)
    skel_true = soft_skel(y_true, iter_)
    pres = ((skel_pred * y_true).sum() + smooth) / (skel_pred.sum() + smooth)
    rec = ((skel_true * y_pred).sum() + smooth) / (skel_true.sum() + smooth)
    return 1.0 - 2.0 * (pres * rec) / (pres + rec)


def soft_dice(y_true, y_pred):
    smooth = 1
    inter = (y_true * y_pred).sum()
    return 1.0 - (2.0 * inter + smooth) / (y_true.sum() + y_pred.sum() + smooth)


def soft_dice_cldice_loss(iters=15, alpha=0.5):
    def loss(y_true, y_pred):
        return (1.0 - alpha) * soft_dice(y_true, y_pred) + alpha * soft_clDice_loss(y_true, y_pred, iters)
    return loss


def min_max_norm(arr, axis=(1, 2, 3, 4)):
    """utils.py:27-48 — no epsilon; gradient flows through min and max (ties split evenly)."""
    mn = arr.amin(dim=axis, keepdim=True)
    mx = arr.amax(dim=axis, keepdim=True)
    return (arr - mn) / (mx - mn)


def reduce_mean(cfg, x, axis=None):
    arr = x.mean() if axis is None else x.mean(dim=axis)
    return arr.sum() / cfg.global_batch_size


def MSE(cfg, y_true, y_pred):
    return reduce_mean(cfg, (y_true - y_pred) ** 2, axis=tuple(range(1, y_true.dim())))


def keras_bce(y_true, y_pred, eps=1e-7):
    """keras.backend.binary_crossentropy(from_logits=False) + mean over the last axis."""
    p = torch.clamp(y_pred, eps, 1.0 - eps)
    bce = y_true * torch.log(p + eps) + (1.0 - y_true) * torch.log(1.0 - p + eps)
    return (-bce).mean(dim=-1)


def gaussian_taps(size=3, sigma=1.5, dtype=torch.float32):
    """loss_functions.py:89-92 — tf.range(-size//2+1, size//2+1) = [-1,0,1] for size 3."""
    grid = torch.arange(-size // 2 + 1, size // 2 + 1, dtype=dtype)
    g = torch.exp(-0.5 * (grid / sigma) ** 2) / (sigma * math.sqrt(2.0 * math.pi))
    return g / g.sum()


def ssim_loss_3d(y_true, y_pred, max_val=1.0, filter_size=3, filter_sigma=1.5, k1=0.01, k2=0.03):
    g = gaussian_taps(filter_size, filter_sigma, y_true.dtype)
    w = torch.einsum("i,j,k->ijk", g, g, g)[None, None]

    def blur(t):  # tf.nn.conv3d(..., padding='SAME'): zero padding
        return _l(F.conv3d(_c(t), w, padding=filter_size // 2))

    mu_t, mu_p = blur(y_true), blur(y_pred)
    mu_tt, mu_pp, mu_tp = mu_t ** 2, mu_p ** 2, mu_t * mu_p
    s_tt = blur(y_true ** 2) - mu_tt
    s_pp = blur(y_pred ** 2) - mu_pp
    s_tp = blur(y_true * y_pred) - mu_tp
    c1, c2 = (k1 * max_val) ** 2, (k2 * max_val) ** 2
    ssim = (2 * mu_tp + c1) * (2 * s_tp + c2) / ((mu_tt + mu_pp + c1) * (s_tt + s_pp + c2))
    return 1.0 - ssim


def cycle_loss(cfg, real, cycled, typ=None):
    if typ == "mse":
        return MSE(cfg, real, cycled) * cfg.lambda_cycle
    if typ is None:
        return reduce_mean(cfg, (real - cycled).abs(), axis=tuple(range(1, real.dim()))) * cfg.lambda_cycle
    # any other string (VanGan passes "bce") takes the else branch, loss_functions.py:185-190
    r = min_max_norm(real)
    c = min_max_norm(cycled)
    return reduce_mean(cfg, keras_bce(r, c)) * cfg.lambda_cycle


def cycle_reconstruction(cfg, real, cycled):
    return reduce_mean(cfg, ssim_loss_3d(min_max_norm(real), min_max_norm(cycled))) * cfg.lambda_reconstruction


def cycle_seg_loss(cfg, real, cycled, iters=15):
    r = min_max_norm(real)
    c = min_max_norm(cycled)
    return soft_dice_cldice_loss(iters=iters)(r, c) * (cfg.lambda_topology / cfg.n_devices)


def generator_loss_fn(cfg, fake):
    return MSE(cfg, torch.ones_like(fake), fake)


def discriminator_loss_fn(cfg, real, fake):
    return 0.5 * (MSE(cfg, torch.ones_like(real), real) + MSE(cfg, torch.zeros_like(fake), fake))
