"""Loss functions with the reference's call signatures (loss_functions.py:7-22,56-68,163-226,
255-322), evaluated by the CUDA kernels in csrc/losses.cu and csrc/skel.cu.

Each function takes the VanGan-like object first (`self` carries global_batch_size, n_devices,
lambda_cycle, lambda_reconstruction, lambda_topology and the per-step LossContext `self.loss_ctx`)
and returns a `Scalar`: `float(s)` is the loss value, `s.seeds()` launches the backward kernels and
returns [(Var, dL/dVar)] for `Tape.backward`.  Scalars add and scale like numbers, so
`total_loss_I = gen_IS_loss + cycle_loss_I + seg_loss` reads as in vangan.py:335.
"""
import torch

from . import engine as E
from ._lib import call
from .clDice_func import cldice_terms


class LossContext:
    """Per-step scratch: one device buffer of double accumulators, read back once."""

    def __init__(self, slots=64):
        self.acc = torch.zeros(slots, dtype=torch.float64, device=E.DEV)
        self.used = 0
        self.host = None
        self._norm = {}

    def slot(self, n=1):
        s = self.used
        self.used += n
        assert self.used <= self.acc.numel()
        self.host = None
        return s

    def ptr(self, s):
        return self.acc.data_ptr() + 8 * s

    def values(self):
        if self.host is None:
            self.host = self.acc.cpu().numpy()   # the one host sync of the loss forward
        return self.host

    def normalized(self, var):
        """min_max_norm_tf(x, axis=(1,2,3,4)) (utils.py:27-48): cached per tensor."""
        key = id(var)
        if key not in self._norm:
            x = var.data
            n = x.shape[0]
            v = x.numel() // n
            mm = torch.empty(2 * n, dtype=torch.float32, device=E.DEV)
            enc_ws = torch.empty(max(2 * n, 64), dtype=torch.int32, device=E.DEV)   # per call: normalisations may run on different streams
            call("vg_minmax", x, n, v, mm, enc_ws)
            nrm = torch.empty_like(x)
            call("vg_minmax_normalize", x, mm, nrm, n, v)
            self._norm[key] = (nrm, mm)
        return self._norm[key]

    def normalize_bwd(self, var, gn):
        """gradient w.r.t. the raw tensor given the gradient w.r.t. its min-max normalised version"""
        nrm, mm = self.normalized(var)
        x = var.data
        n = x.shape[0]
        dx = torch.empty_like(x)
        ws = torch.empty(4 * n, dtype=torch.float64, device=E.DEV)
        call("vg_minmax_normalize_bwd", x, nrm, mm, gn, dx, n, x.numel() // n, ws, 0)
        return dx


class Scalar:
    def __init__(self, ctx, value_fn, grad_fns):
        self.ctx, self.value_fn, self.grad_fns = ctx, value_fn, grad_fns

    def __add__(self, o):
        if isinstance(o, (int, float)):
            return Scalar(self.ctx, lambda a: self.value_fn(a) + o, self.grad_fns)
        return Scalar(self.ctx, lambda a: self.value_fn(a) + o.value_fn(a), self.grad_fns + o.grad_fns)

    __radd__ = __add__

    def __mul__(self, c):
        c = float(c)
        return Scalar(self.ctx, lambda a: self.value_fn(a) * c,
                      [(lambda a, s, f=f: f(a, s * c)) for f in self.grad_fns])

    __rmul__ = __mul__

    def part(self, *idx):
        """The same scalar restricted (for the BACKWARD pass only) to the gradient terms idx of its sum: backward is linear in the
        seeds, so a loss may be swept as several independent parts (vangan._sweeps runs them on separate streams)."""
        return Scalar(self.ctx, self.value_fn, [self.grad_fns[i] for i in idx])

    def __float__(self):
        return float(self.value_fn(self.ctx.values()))

    def numpy(self):
        return float(self)

    def seeds(self, scale=1.0):
        """Launches the backward kernels of every term.  No host synchronisation: coefficients that depend on reduced sums
        are computed on the device (vg_cldice_coeffs), so `a` (the host copy of the accumulators) is not needed."""
        a = None
        out = []
        for f in self.grad_fns:
            out += f(a, scale)
        return out


def _data(x):
    return x.data if isinstance(x, E.Var) else x


def reduce_mean(self, inputs_sum, count_per_sample, n_local, axis=None):
    """loss_functions.py:7-22 on an already reduced SUM: axis=None -> mean over the whole local batch
    tensor, else mean per sample then sum over the local batch; both divided by the GLOBAL batch size."""
    if axis is None:
        return inputs_sum / (count_per_sample * n_local) / self.global_batch_size
    return inputs_sum / count_per_sample / self.global_batch_size


def MSE(self, y_true, y_pred):
    """loss_functions.py:56-68.  y_true: Var/tensor or a python constant (ones_like / zeros_like)."""
    ctx = self.loss_ctx
    p = _data(y_pred)
    n = p.shape[0]
    per = p.numel() // n
    const = isinstance(y_true, (int, float))
    t = None if const else _data(y_true)
    s = ctx.slot()
    call("vg_sqdiff_sum", p, t, float(y_true) if const else 0.0, p.numel(), ctx.ptr(s))
    G = self.global_batch_size

    def value(a):
        return a[s] / per / G

    def grad(a, scale):
        if not isinstance(y_pred, E.Var):
            return []
        k = 2.0 * scale / (per * G)
        g = torch.empty_like(p)
        if const:
            call("vg_lincomb", g, p.numel(), 0, -k * float(y_true), p, k, None, 0.0, None, 0.0)
        else:
            call("vg_lincomb", g, p.numel(), 0, 0.0, p, k, t, -k, None, 0.0)
        return [(y_pred, g)]

    return Scalar(ctx, value, [grad])


def cycle_loss(self, real_image, cycled_image, typ=None):
    """loss_functions.py:163-190.  'mse' -> MSE; any other string (VanGan passes "bce") -> per-sample
    min-max norm + Keras BinaryCrossentropy, mean over the local batch tensor."""
    if typ == "mse":
        return MSE(self, real_image, cycled_image) * self.lambda_cycle
    if typ is None or typ == "L4":
        raise NotImplementedError("cycle_loss: MAE / L4 branches are not reached by VanGan.compute_losses")
    ctx = self.loss_ctx
    nr, _ = ctx.normalized(real_image)
    nc, _ = ctx.normalized(cycled_image)
    n = nc.shape[0]
    per = nc.numel() // n
    s = ctx.slot()
    call("vg_bce_sum", nr, nc, nc.numel(), ctx.ptr(s))
    G, lam = self.global_batch_size, self.lambda_cycle

    def value(a):
        return a[s] / (per * n) / G * lam

    def grad(a, scale):
        gn = torch.empty_like(nc)
        call("vg_bce_bwd", nr, nc, scale * lam / (per * n * G), gn, nc.numel(), 0)
        return [(cycled_image, ctx.normalize_bwd(cycled_image, gn))]

    return Scalar(ctx, value, [grad])


def cycle_reconstruction(self, real_image, cycled_image):
    """loss_functions.py:193-208 + ssim_loss_3d (:86-117)."""
    ctx = self.loss_ctx
    nr, _ = ctx.normalized(real_image)
    nc, _ = ctx.normalized(cycled_image)
    n, d, h, w, _c = nc.shape
    per = d * h * w
    s = ctx.slot()
    mA, mB, mC = torch.empty_like(nc), torch.empty_like(nc), torch.empty_like(nc)
    call("vg_ssim_fwd", nr, nc, n, d, h, w, ctx.ptr(s), mA, mB, mC)
    G, lam = self.global_batch_size, self.lambda_reconstruction

    def value(a):
        return a[s] / (per * n) / G * lam

    def grad(a, scale):
        gn = torch.empty_like(nc)
        call("vg_ssim_bwd", nr, nc, mA, mB, mC, n, d, h, w, scale * lam / (per * n * G), gn, 0)
        return [(cycled_image, ctx.normalize_bwd(cycled_image, gn))]

    return Scalar(ctx, value, [grad])


def cycle_seg_loss(self, real_image, cycled_image, iters=15, alpha=0.5):
    """loss_functions.py:211-226: soft_dice_cldice_loss()(minmax(real), minmax(cycled)) * lambda_topology / n_devices."""
    ctx = self.loss_ctx
    nr, _ = ctx.normalized(real_image)
    nc, _ = ctx.normalized(cycled_image)
    value, grad_n = cldice_terms(ctx, nr, nc, iters, alpha, self.lambda_topology / self.n_devices)

    def grad(a, scale):
        return [(cycled_image, ctx.normalize_bwd(cycled_image, grad_n(a, scale)))]

    return Scalar(ctx, value, [grad])


def generator_loss_fn(self, fake_image, typ=None, from_logits=True):
    """loss_functions.py:255-286, LSGAN branch (typ=None): MSE(ones, D(fake))."""
    if typ is not None:
        raise NotImplementedError("generator_loss_fn: bce/bfce branches are not reached by VanGan.compute_losses")
    return MSE(self, 1.0, fake_image)


def discriminator_loss_fn(self, real_image, fake_image, typ=None, from_logits=True):
    """loss_functions.py:289-322, LSGAN branch: 0.5*(MSE(ones, D(real)) + MSE(zeros, D(fake)))."""
    if typ is not None:
        raise NotImplementedError("discriminator_loss_fn: bce/bfce branches are not reached by VanGan.compute_losses")
    return (MSE(self, 1.0, real_image) + MSE(self, 0.0, fake_image)) * 0.5
