"""clDice soft-skeleton and soft-Dice/clDice loss (reference clDice_func.py:8-149) on the fused
CUDA stencil kernels of csrc/skel.cu.  3-D branch only (DIMENSIONS=3 in main.py:80).
"""
import torch

from . import engine as E
from ._lib import call, lib


def _skel_buffers(x, iters):
    n, d, h, w = x.shape[:4]
    nv = x.numel()
    Eb = torch.empty((iters + 2, nv), dtype=torch.float32, device=E.DEV)
    Sb = torch.empty((iters + 1, nv), dtype=torch.float32, device=E.DEV)
    call("vg_soft_skel_fwd", x, Eb, Sb, n, d, h, w, iters, work=16.0 * nv * (iters + 1))   # HBM model 16 B * V * (k+1)
    return Eb, Sb


def soft_skel(img, iters):
    """soft_skel(img, iters) (clDice_func.py:60-80).  img: (N,D,H,W,1) fp32 CUDA tensor."""
    img = img.contiguous()
    assert img.dim() == 5 and img.shape[-1] == 1 and img.dtype == torch.float32
    _, Sb = _skel_buffers(img, iters)
    return Sb[iters].view(img.shape).clone()


def soft_erode(img):
    """soft_erode (clDice_func.py:8-26): first level of the erosion pyramid."""
    img = img.contiguous()
    Eb, _ = _skel_buffers(img, 0)
    return Eb[1].view(img.shape).clone()


def soft_skel_with_grad(img, iters):
    """Returns (skel, backward) where backward(gskel) -> d/d img."""
    img = img.contiguous()
    n, d, h, w = img.shape[:4]
    Eb, Sb = _skel_buffers(img, iters)

    def backward(gskel):
        nb = lib().vg_soft_skel_bwd_workspace_bytes(n, d, h, w)
        ws = torch.empty(nb // 4, dtype=torch.float32, device=E.DEV)
        dx = torch.empty_like(img)
        call("vg_soft_skel_bwd", Eb, Sb, gskel.contiguous(), dx, ws, nb, n, d, h, w, iters,
             # HBM bytes of this formulation per voxel and level: coefficient pass reads G, S_{j-1}, E_j, E_{j+1} and writes a_j, G_{j-1};
             # routing pass reads E_j, a_j, D_{j+1}, a_{j-1} and writes D_j -> 11 floats
             work=44.0 * img.numel() * (iters + 1))
        return dx

    return Sb[iters].view(img.shape), backward


def cldice_terms(ctx, y_true, y_pred, iters, alpha, scale0):
    """(1-alpha)*soft_dice + alpha*soft_clDice_loss (clDice_func.py:83-149) times scale0, as
    (value_fn(acc), grad_fn(acc, scale) -> dL/d y_pred).  Sums run over the whole local batch tensor."""
    skel_p, skel_bwd = soft_skel_with_grad(y_pred, iters)
    skel_t = soft_skel(y_true, iters)
    s = ctx.slot(7)
    call("vg_cldice_sums", y_true, y_pred, skel_t, skel_p, y_pred.numel(), ctx.ptr(s))
    smooth = 1.0

    def parts(a):
        s0, s1, s2, s3, s4, s5, s6 = [float(a[s + i]) for i in range(7)]
        P = (s0 + smooth) / (s1 + smooth)
        R = (s2 + smooth) / (s3 + smooth)
        cl = 1.0 - 2.0 * (P * R) / (P + R)
        den = s5 + s6 + smooth
        dice = 1.0 - (2.0 * s4 + smooth) / den
        return s0, s1, s2, s3, s4, den, P, R, cl, dice

    def value(a):
        *_, cl, dice = parts(a)
        return ((1.0 - alpha) * dice + alpha * cl) * scale0

    def grad(a, scale):
        # the coefficients depend on the seven sums: computed on the device, so the backward needs no host round trip
        coef = torch.empty(8, dtype=torch.float32, device=E.DEV)
        call("vg_cldice_coeffs", ctx.ptr(s), float(alpha), float(scale * scale0), coef)
        # seed for the skeleton backward: d loss / d skel_pred = kc * dcl/dP * dP/dskel
        gsk = torch.empty_like(y_pred)
        call("vg_lincomb_dev", gsk, gsk.numel(), 0, coef, y_true, None, None)
        d0 = skel_bwd(gsk)
        gn = torch.empty_like(y_pred)
        call("vg_lincomb_dev", gn, gn.numel(), 0, coef[2:], y_true, skel_t, d0)
        return gn

    return value, grad


def soft_dice_cldice_loss(iters=15, alpha=0.5):
    """clDice_func.py:122-149: returns loss(y_true, y_pred) -> python float (forward value)."""
    def loss(y_true, y_pred):
        from .loss_functions import LossContext
        ctx = LossContext()
        value, _ = cldice_terms(ctx, y_true.contiguous(), y_pred.contiguous(), iters, alpha, 1.0)
        return value(ctx.values())
    return loss
