"""3D ResNet generator builder — same call signature as the reference's `get_resnet_generator` (generator.py:7-73), `VanGan`'s
default `gen_i2s` / `gen_s2i` = 'resnet' (vangan.py:29-30,88-95,126-133: num_downsampling_blocks = num_upsample_blocks = 3), backed by the
C-ABI kernels.

    ReflectionPadding3D(1) -> Conv3D(32, 7, valid, no bias) -> InstanceNorm -> ReLU -> SpatialDropout3D(0.5)          generator.py:35-42
    3 x downsample: ReflectionPadding3D(1) -> Conv3D(k3, s2, valid, no bias) -> InstanceNorm -> ReLU -> SpatialDropout3D(0.2)
                                                                                          generator.py:45-50, building_blocks.py:126-196
    6 x residual_block: pad -> Conv3D k3 -> IN -> ReLU -> pad -> Conv3D k3 -> IN -> add    generator.py:53-56, building_blocks.py:68-123
    3 x upsample: UpSampling3D(2) -> Conv3D(k4, s1, 'same', no bias) -> InstanceNorm -> ReLU           generator.py:59-63, :240-280
    Conv3D(1, 7, 'same') -> tanh                                                                        generator.py:66-69

A 128^3 input becomes 124^3 x 32 after the valid 7^3 convolution, 62 / 31 / 16 after the stride-2 blocks and 32 / 64 / 128 on the way
up, so the output is 128^3 again.  The 7^3 convolutions have one single-channel side and run as direct kernels (csrc/conv_k7.cu); the
padding each convolution needs (reflect 1, TF-'same' zeros 1/2 for k4, zeros 3/3 for k7) is written by the pass that produces its
input.  The `ReflectionPadding3D(2)` of generator.py:66-67 belongs to num_downsampling_blocks == 2 and is not built.
"""
from collections import OrderedDict

import torch

from . import engine as E
from ._lib import ACT_NONE, ACT_RELU, ACT_TANH, PAD_REFLECT, PAD_ZERO


def resnet_param_shapes(filters=32, num_downsampling_blocks=3, num_residual_blocks=6, num_upsample_blocks=3, cin=1):
    P = OrderedDict()

    def inorm(name, c):
        P[name + ".gamma"] = (c,); P[name + ".beta"] = (c,)

    f = filters
    P["c0.conv.w"] = (7, 7, 7, cin, f); inorm("c0.in", f)
    for i in range(num_downsampling_blocks):
        P["down%d.conv.w" % i] = (3, 3, 3, f, 2 * f); inorm("down%d.in" % i, 2 * f)
        f *= 2
    for j in range(num_residual_blocks):
        for k in (1, 2):
            P["res%d.c%d.conv.w" % (j, k)] = (3, 3, 3, f, f); inorm("res%d.c%d.in" % (j, k), f)
    for i in range(num_upsample_blocks):
        P["up%d.conv.w" % i] = (4, 4, 4, f, f // 2); inorm("up%d.in" % i, f // 2)
        f //= 2
    P["out.conv.w"] = (7, 7, 7, f, 1); P["out.conv.b"] = (1,)
    return P


class ResNetGenerator(E.Network):
    def __init__(self, name, filters=32, num_downsampling_blocks=3, num_residual_blocks=6, num_upsample_blocks=3, cin=1, seed=None):
        super().__init__(name, resnet_param_shapes(filters, num_downsampling_blocks, num_residual_blocks, num_upsample_blocks, cin))
        self.nd, self.nr, self.nu, self.filters = num_downsampling_blocks, num_residual_blocks, num_upsample_blocks, filters
        f = filters
        self.conv0 = E.Conv3D(self, "c0.conv", 7, 1, cin, f, use_bias=False)
        self.norm0 = E.InstanceNorm(self, "c0.in", f)
        self.down = []
        for i in range(self.nd):
            self.down.append((E.Conv3D(self, "down%d.conv" % i, 3, 2, f, 2 * f, use_bias=False), E.InstanceNorm(self, "down%d.in" % i, 2 * f)))
            f *= 2
        self.res = []
        for j in range(self.nr):
            self.res.append((E.Conv3D(self, "res%d.c1.conv" % j, 3, 1, f, f, use_bias=False), E.InstanceNorm(self, "res%d.c1.in" % j, f),
                             E.Conv3D(self, "res%d.c2.conv" % j, 3, 1, f, f, use_bias=False), E.InstanceNorm(self, "res%d.c2.in" % j, f)))
        self.up = []
        for i in range(self.nu):
            self.up.append((E.Conv3D(self, "up%d.conv" % i, 4, 1, f, f // 2, use_bias=False, dx_crop=(1, 2)), E.InstanceNorm(self, "up%d.in" % i, f // 2)))
            f //= 2
        self.out = E.Conv3D(self, "out.conv", 7, 1, f, 1, act=ACT_TANH, dx_crop=(3, 3))
        self.rng_step = 0
        if seed is not None:
            # kernel_initializer = gamma_initializer = 'he_normal' (generator.py:14-15); the final Conv3D keeps Keras' glorot_uniform
            self.load(E.default_init({n: p.shape for n, p in self.params.items()}, seed, glorot=("out.conv.w",), he_gamma=True))

    def drop_channels(self):
        """channel widths of the SpatialDropout3D layers in call order: 0.5 after the first block, 0.2 after every downsample block"""
        return [self.filters * 2 ** i for i in range(self.nd + 1)]

    def drop_rates(self):
        return [0.5] + [0.2] * self.nd

    def forward(self, tape, x, training=True, masks=None, seed=0, seed_dev=None, taps=None):
        """x: Var (N,D,H,W,1) fp32.  masks: optional explicit SpatialDropout3D masks [(N, C) scaled by 1/(1-rate)] in call order; None ->
        drawn on the device (Philox keyed on seed + *seed_dev) when training."""
        n = x.shape[0]
        widths, rates = self.drop_channels(), self.drop_rates()

        def mask(i):
            if not training:
                return None
            if masks is not None:
                return masks[i].reshape(n * widths[i]).to(torch.float32).contiguous()
            return E.dropout_mask(n * widths[i], rates[i], seed * 16 + i, seed_dev)

        h = self.conv0(tape, E.pad_noise(tape, x))
        h = self.norm0(tape, h, act=ACT_RELU, drop=mask(0), pad=(1, 1, PAD_REFLECT))
        if taps is not None:
            taps["c0"] = h
        for i, (conv, norm) in enumerate(self.down):
            last = i == self.nd - 1
            h = norm(tape, conv(tape, h), act=ACT_RELU, drop=mask(i + 1), pad=(0, 0, PAD_ZERO) if last else (1, 1, PAD_REFLECT))
            if taps is not None:
                taps["down%d" % i] = h
        for j, (c1, n1, c2, n2) in enumerate(self.res):
            p = E.gather_pad(tape, None, h, up=1, pad=1, mode=PAD_REFLECT)
            c = n1(tape, c1(tape, p), act=ACT_RELU, pad=(1, 1, PAD_REFLECT))
            h = n2(tape, c2(tape, c), act=ACT_NONE, residual=h)
            if taps is not None:
                taps["res%d" % j] = h
        for i, (conv, norm) in enumerate(self.up):
            last = i == self.nu - 1
            u = E.upsample_pad(tape, h, 1, 2)                       # UpSampling3D(2) + TF 'same' zeros of the k4 convolution
            h = norm(tape, conv(tape, u), act=ACT_RELU, pad=(3, 3, PAD_ZERO) if last else (0, 0, PAD_ZERO))   # zeros 3/3: the k7 'same' head
            if taps is not None:
                taps["up%d" % i] = h
        return self.out(tape, h)

    def __call__(self, x, training=False):
        xt = torch.as_tensor(x, dtype=torch.float32, device=E.DEV).contiguous()
        return self.forward(E.Tape(enabled=False), E.Var(xt), training=training, seed=self.rng_step).data


def get_resnet_generator(input_img_size=(64, 64, 512, 1), batch_size=None, filters=32, num_downsampling_blocks=2, num_residual_blocks=6,
                         num_upsample_blocks=2, gamma_initializer='he_normal', kernel_initializer='he_normal', name=None, seed=0):
    """Same arguments as the reference builder (generator.py:7-17)."""
    if num_downsampling_blocks != num_upsample_blocks:
        raise ValueError("num_downsampling_blocks and num_upsample_blocks must match for the output to have the input's size")
    if num_downsampling_blocks == 2:
        raise NotImplementedError("get_resnet_generator: the ReflectionPadding3D(2) head of num_downsampling_blocks == 2 "
                                  "(generator.py:66-67) is not built; VanGan uses 3 (vangan.py:93-94)")
    if gamma_initializer != 'he_normal' or kernel_initializer != 'he_normal':
        raise NotImplementedError("get_resnet_generator: initializers other than the reference's defaults")
    return ResNetGenerator(name or "generator", filters=filters, num_downsampling_blocks=num_downsampling_blocks,
                           num_residual_blocks=num_residual_blocks, num_upsample_blocks=num_upsample_blocks, cin=input_img_size[-1], seed=seed)
