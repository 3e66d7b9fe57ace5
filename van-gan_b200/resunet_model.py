"""ResUNet generator builder — same call signature as the reference's `ResUNet`
(resunet_model.py:185-249), backed by the C-ABI kernels instead of Keras layers.

Layer order, padding and parameter layouts follow resunet_model.py:23-182 with the options VanGan
uses (vangan.py:112-122): upsample_mode='simple' (UpSampling3D + concatenate), dropout_type='none',
no input noise, no attention gate, tanh head.  Unsupported options raise (no silent fallback).
"""
from collections import OrderedDict

from . import engine as E
from ._lib import ACT_NONE, ACT_RELU, ACT_TANH, PAD_REFLECT, PAD_ZERO


def resunet_param_shapes(filters=16, num_layers=4, cin=1):
    f = [filters * (2 ** i) for i in range(num_layers + 1)]
    P = OrderedDict()

    def conv(name, k, ci, co):
        P[name + ".w"] = (k, k, k, ci, co)
        P[name + ".b"] = (co,)

    def inorm(name, c):
        P[name + ".gamma"] = (c,)
        P[name + ".beta"] = (c,)

    def resblock(name, ci, co):
        inorm(name + ".cb1.in", ci); conv(name + ".cb1.conv", 3, ci, co)
        inorm(name + ".cb2.in", co); conv(name + ".cb2.conv", 3, co, co)
        conv(name + ".short.conv", 1, ci, co); inorm(name + ".short.in", co)

    conv("stem.conv0", 3, cin, f[0])
    inorm("stem.cb.in", f[0]); conv("stem.cb.conv", 3, f[0], f[0])
    conv("stem.short.conv", 1, cin, f[0]); inorm("stem.short.in", f[0])
    for e in range(1, num_layers + 1):
        resblock("enc%d" % e, f[e - 1], f[e])
    for i in (1, 2):
        inorm("bridge%d.in" % i, f[-1]); conv("bridge%d.conv" % i, 3, f[-1], f[-1])
    for d in reversed(range(num_layers)):
        resblock("dec%d" % d, f[d + 1] + f[d], f[d])
    conv("head", 1, f[0], 1)
    return P


class _ConvBlock:
    """conv_block (resunet_model.py:42-66): InstanceNorm -> ReLU -> ReflectionPadding3D -> Conv3D."""

    def __init__(self, net, name, ci, co, stride):
        self.norm = E.InstanceNorm(net, name + ".in", ci)
        self.conv = E.Conv3D(net, name + ".conv", 3, stride, ci, co)

    def __call__(self, tape, x):
        return self.conv(tape, self.norm(tape, x, act=ACT_RELU, pad=(1, 1, PAD_REFLECT)))


class _ResBlock:
    """residual_block (resunet_model.py:103-143)."""

    def __init__(self, net, name, ci, co, stride):
        self.cb1 = _ConvBlock(net, name + ".cb1", ci, co, stride)
        self.cb2 = _ConvBlock(net, name + ".cb2", co, co, 1)
        self.short = E.Conv3D(net, name + ".short.conv", 1, stride, ci, co)
        self.short_norm = E.InstanceNorm(net, name + ".short.in", co)

    def __call__(self, tape, x):
        res = self.cb2(tape, self.cb1(tape, x))
        sc = self.short(tape, x)
        return self.short_norm(tape, sc, act=ACT_NONE, residual=res)   # Add()([shortcut, res])


class ResUNetModel(E.Network):
    def __init__(self, name, filters=16, num_layers=4, cin=1, seed=None):
        super().__init__(name, resunet_param_shapes(filters, num_layers, cin))
        f = [filters * (2 ** i) for i in range(num_layers + 1)]
        self.num_layers = num_layers
        self.stem_conv0 = E.Conv3D(self, "stem.conv0", 3, 1, cin, f[0])
        self.stem_cb = _ConvBlock(self, "stem.cb", f[0], f[0], 1)
        self.stem_short = E.Conv3D(self, "stem.short.conv", 1, 1, cin, f[0])
        self.stem_short_norm = E.InstanceNorm(self, "stem.short.in", f[0])
        self.enc = [_ResBlock(self, "enc%d" % e, f[e - 1], f[e], 2) for e in range(1, num_layers + 1)]
        self.bridge = [_ConvBlock(self, "bridge%d" % i, f[-1], f[-1], 1) for i in (1, 2)]
        self.dec = {d: _ResBlock(self, "dec%d" % d, f[d + 1] + f[d], f[d], 1) for d in range(num_layers)}
        self.head = E.Conv3D(self, "head", 1, 1, f[0], 1, act=ACT_TANH)
        if seed is not None:
            self.load(E.default_init({n: p.shape for n, p in self.params.items()}, seed, glorot=("stem.conv0.w", "stem.short.conv.w", "head.w")))

    def forward(self, tape, x, taps=None):
        """x: Var holding an (N,D,H,W,1) fp32 volume.  Returns the (N,D,H,W,1) fp32 tanh output.
        `taps`: optional dict receiving the block outputs (layer-wise parity tests)."""
        conv = self.stem_conv0(tape, E.pad_noise(tape, x))           # stem(), resunet_model.py:87-91
        conv = self.stem_cb(tape, conv)
        sc = self.stem_short(tape, x)
        h = self.stem_short_norm(tape, sc, act=ACT_NONE, residual=conv)
        skips = [h]
        if taps is not None:
            taps["stem"] = h
        for e, blk in enumerate(self.enc):
            h = blk(tape, h)
            skips.append(h)
            if taps is not None:
                taps["enc%d" % (e + 1)] = h
        for blk in self.bridge:
            h = blk(tape, h)
        if taps is not None:
            taps["bridge"] = h
        for d in reversed(range(self.num_layers)):
            h = E.upsample_concat(tape, h, skips[d])
            h = self.dec[d](tape, h)
            if taps is not None:
                taps["dec%d" % d] = h
        return self.head(tape, h)

    def __call__(self, x, training=False):
        """Keras-style call on a torch/numpy NDHWC array; returns a torch CUDA fp32 tensor."""
        import torch
        xt = torch.as_tensor(x, dtype=torch.float32, device=E.DEV).contiguous()
        return self.forward(E.Tape(enabled=False), E.Var(xt)).data


def ResUNet(input_shape, upsample_mode='deconv', dropout=0.2, dropout_change_per_layer=0.0, dropout_type='none',
            kernel_initializer='he_normal', use_attention_gate=False, filters=16, num_layers=4,
            output_activation='tanh', use_input_noise=False, name='resunet', seed=0):
    """Same arguments as the reference builder (resunet_model.py:185-197).  Only the combination the
    reference's VanGan instantiates is implemented on the CUDA path; anything else raises."""
    if upsample_mode == 'deconv' or use_attention_gate or use_input_noise or dropout_type not in ('none', None):
        raise NotImplementedError("ResUNet: only upsample_mode='simple', dropout_type='none', no attention gate, "
                                  "no input noise is built (vangan.py:112-122)")
    if output_activation != 'tanh' or kernel_initializer != 'he_normal':
        raise NotImplementedError("ResUNet: tanh head / he_normal only")
    return ResUNetModel(name, filters=filters, num_layers=num_layers, cin=input_shape[-1], seed=seed)
