"""V-Net generator builder — same call signature as the reference's `custom_vnet`
(vnet_model.py:149-268), backed by the C-ABI kernels instead of Keras layers.

Built: both variants VanGan instantiates.  gen_IS (vangan.py:97-110): InstanceNormalization (`use_batch_norm=False`, so every
Conv3D has a bias), `upsample_mode='upsample'` (UpSampling3D + Conv3D k3 'same').  gen_SI (vangan.py:135-149):
BatchNormalization (block convolutions without bias, per-replica batch statistics, Keras moving averages for inference),
`upsample_mode='deconv'` (Conv3DTranspose k2 s2 = pointwise GEMM + depth-to-space).  Both: SpatialDropout3D(0.5) after the first
norm of every encoder / bottleneck block, no dropout on the decoder, no attention gate, tanh head.  Layer order follows `conv3d_block` (vnet_model.py:80-146):
pad -> Conv3D(relu) -> norm -> [dropout] -> pad -> Conv3D(relu) -> norm.  The ReLU that Keras applies inside the
convolution is applied by the InstanceNorm kernels on load (`relu_input`), so the convolutions stay linear.

Not built (raise): attention gate, `addnoise`, 'standard' dropout.
"""
from collections import OrderedDict

import torch

from . import engine as E
from ._lib import ACT_NONE, ACT_TANH, PAD_REFLECT, PAD_ZERO


def vnet_param_shapes(filters=32, num_layers=4, cin=1, use_batch_norm=False, upsample_mode='upsample'):
    """Trainable variables in Keras creation order.  use_batch_norm: the block convolutions have no bias (vnet_model.py:124,139) and
    are followed by BatchNormalization (gamma, beta trainable; moving statistics are buffers); upsample_mode 'deconv': the decoder
    upsampling is a Conv3DTranspose with kernel (2,2,2,Cout,Cin) + bias (vnet_model.py:245)."""
    P = OrderedDict()
    nk = "bn" if use_batch_norm else "in"

    def conv(name, k, ci, co, bias=True):
        P[name + ".w"] = (k, k, k, ci, co)
        if bias:
            P[name + ".b"] = (co,)

    def block(name, ci, co):
        for j, c_in in ((1, ci), (2, co)):
            conv("%s.c%d.conv" % (name, j), 3, c_in, co, bias=not use_batch_norm)
            P["%s.c%d.%s.gamma" % (name, j, nk)] = (co,)
            P["%s.c%d.%s.beta" % (name, j, nk)] = (co,)

    f, ci = filters, cin
    for l in range(num_layers):
        block("enc%d" % l, ci, f)
        ci, f = f, f * 2
    block("bridge", ci, f)
    for l in reversed(range(num_layers)):
        f //= 2
        if upsample_mode == 'deconv':
            P["dec%d.up.w" % l] = (2, 2, 2, f, 2 * f)
            P["dec%d.up.b" % l] = (f,)
        else:
            conv("dec%d.up.conv" % l, 3, 2 * f, f)
        block("dec%d" % l, 2 * f, f)
    conv("head", 1, f, 1)
    return P


class _Block:
    """conv3d_block (vnet_model.py:80-146) over an input that is already reflect-padded."""

    def __init__(self, net, name, ci, co, use_batch_norm=False):
        self.bn = use_batch_norm
        Norm = E.BatchNorm if use_batch_norm else E.InstanceNorm
        nk = "bn" if use_batch_norm else "in"
        self.conv1 = E.Conv3D(net, name + ".c1.conv", 3, 1, ci, co, use_bias=not use_batch_norm)
        self.norm1 = Norm(net, "%s.c1.%s" % (name, nk), co)
        self.conv2 = E.Conv3D(net, name + ".c2.conv", 3, 1, co, co, use_bias=not use_batch_norm)
        self.norm2 = Norm(net, "%s.c2.%s" % (name, nk), co)

    def __call__(self, tape, xpad, drop=None, training=True):
        kw = dict(training=training) if self.bn else {}
        c = self.conv1(tape, xpad)
        c = self.norm1(tape, c, act=ACT_NONE, drop=drop, pad=(1, 1, PAD_REFLECT), relu_input=True, **kw)
        c = self.conv2(tape, c)
        return self.norm2(tape, c, act=ACT_NONE, relu_input=True, **kw)


class VNetModel(E.Network):
    def __init__(self, name, filters=32, num_layers=4, cin=1, dropout=0.5, seed=None, use_batch_norm=False, upsample_mode='upsample'):
        super().__init__(name, vnet_param_shapes(filters, num_layers, cin, use_batch_norm, upsample_mode))
        self.num_layers, self.rate, self.filters = num_layers, dropout, filters
        self.use_batch_norm, self.upsample_mode = use_batch_norm, upsample_mode
        f, ci = filters, cin
        self.enc = []
        for l in range(num_layers):
            self.enc.append(_Block(self, "enc%d" % l, ci, f, use_batch_norm))
            ci, f = f, f * 2
        self.bridge = _Block(self, "bridge", ci, f, use_batch_norm)
        self.up, self.dec = {}, {}
        for l in reversed(range(num_layers)):
            f //= 2
            if upsample_mode == 'deconv':
                self.up[l] = E.Conv3DTranspose(self, "dec%d.up" % l, 2 * f, f)
            else:
                self.up[l] = E.Conv3D(self, "dec%d.up.conv" % l, 3, 1, 2 * f, f, dx_crop=(1, 1))   # k3 'same': zero pad 1/1
            self.dec[l] = _Block(self, "dec%d" % l, 2 * f, f, use_batch_norm)
        self.head = E.Conv3D(self, "head", 1, 1, f, 1, act=ACT_TANH)
        self.rng_step = 0
        if seed is not None:
            self.load(E.default_init({n: p.shape for n, p in self.params.items()}, seed, glorot=("head.w",) + tuple(n for n in self.params if n.endswith(".up.w"))))

    def drop_channels(self):
        """channel widths of the SpatialDropout3D layers, in call order (encoder blocks, then the bottleneck)."""
        return [self.filters * (2 ** l) for l in range(self.num_layers + 1)]

    def forward(self, tape, x, training=True, masks=None, seed=0, taps=None):
        """x: Var (N,D,H,W,1) fp32.  masks: optional explicit dropout masks, one (N, C) tensor per encoder /
        bottleneck block (already scaled by 1/(1-rate)); None -> drawn on the device from `seed` when training."""
        n = x.shape[0]

        def mask(i, c):
            if not training or self.rate <= 0.0:
                return None
            if masks is not None:
                return masks[i].reshape(n * c).to(torch.float32).contiguous()
            g = torch.Generator(device=E.DEV)
            g.manual_seed(seed * 11 + i)
            keep = torch.rand(n * c, device=E.DEV, generator=g) >= self.rate
            return keep.to(torch.float32) / (1.0 - self.rate)

        widths = self.drop_channels()
        h = E.pad_noise(tape, x)                                         # ReflectionPadding3D of the fp32 input
        skips = []
        for l, blk in enumerate(self.enc):
            h = blk(tape, h, drop=mask(l, widths[l]), training=training)
            skips.append(h)
            if taps is not None:
                taps["enc%d" % l] = h
            h = E.maxpool_pad(tape, h, pad=1, mode=PAD_REFLECT)          # MaxPooling3D(2) + next block's pad
        h = self.bridge(tape, h, drop=mask(self.num_layers, widths[-1]), training=training)
        if taps is not None:
            taps["bridge"] = h
        for l in reversed(range(self.num_layers)):
            if self.upsample_mode == 'deconv':
                u = self.up[l](tape, h)                                      # Conv3DTranspose k2 s2
            else:
                u = E.gather_pad(tape, h, None, up=2, pad=1, mode=PAD_ZERO)  # UpSampling3D(2) + 'same' zeros
                u = self.up[l](tape, u)
            h = E.gather_pad(tape, u, skips[l], up=1, pad=1, mode=PAD_REFLECT)   # concatenate([x, conv]) + pad
            h = self.dec[l](tape, h, training=training)
            if taps is not None:
                taps["dec%d" % l] = h
        return self.head(tape, h)

    def __call__(self, x, training=False):
        xt = torch.as_tensor(x, dtype=torch.float32, device=E.DEV).contiguous()
        return self.forward(E.Tape(enabled=False), E.Var(xt), training=training, seed=self.rng_step).data


def custom_vnet(input_shape, num_classes=1, activation='relu', use_batch_norm=True, upsample_mode='deconv', dropout=0.5,
                dropout_change_per_layer=0.0, dropout_type='spatial', use_dropout_on_upsampling=False,
                kernel_initializer='he_normal', use_attention_gate=False, filters=16, num_layers=4,
                output_activation='sigmoid', addnoise=False, name='vnet', seed=0):
    """Same arguments as the reference builder (vnet_model.py:149-165)."""
    if upsample_mode not in ('deconv', 'upsample', 'simple'):
        raise ValueError("upsample_mode must be 'deconv' or 'upsample'")
    if (use_attention_gate or addnoise or dropout_type != 'spatial' or use_dropout_on_upsampling or dropout_change_per_layer != 0.0
            or activation != 'relu' or output_activation != 'tanh' or num_classes != 1 or kernel_initializer != 'he_normal'):
        raise NotImplementedError("custom_vnet: only the option sets VanGan uses (vangan.py:97-110 gen_IS, :135-149 gen_SI) are built")
    return VNetModel(name, filters=filters, num_layers=num_layers, cin=input_shape[-1], dropout=dropout, seed=seed,
                     use_batch_norm=use_batch_norm, upsample_mode='deconv' if upsample_mode == 'deconv' else 'upsample')
