"""3D PatchGAN discriminator builder — same call signature as the reference's `get_discriminator`
(discriminator.py:7-124), with `downsample` from building_blocks.py:126-196 folded in.

With VanGan's arguments (vangan.py:167-192): filters 64, k4 convs, InstanceNorm after the first four
convs, LeakyReLU(0.2), GaussianNoise(0.1) before every conv, SpatialDropout3D(0.2) after blocks 1-3.
The SpectralNormalization and Wasserstein branches are not built (never enabled by the reference).
"""
from collections import OrderedDict

import torch

from . import engine as E
from ._lib import ACT_LEAKY, PAD_REFLECT, PAD_ZERO


def disc_param_shapes(filters=64, cin=1):
    P = OrderedDict()
    P["d0.conv.w"] = (4, 4, 4, cin, filters); P["d0.conv.b"] = (filters,)
    P["d0.in.gamma"] = (filters,); P["d0.in.beta"] = (filters,)
    c = filters
    for i in (1, 2, 3):
        P["d%d.conv.w" % i] = (4, 4, 4, c, 2 * c)
        P["d%d.in.gamma" % i] = (2 * c,); P["d%d.in.beta" % i] = (2 * c,)
        c *= 2
    P["dout.conv.w"] = (3, 3, 3, c, 1); P["dout.conv.b"] = (1,)
    return P


class DiscriminatorModel(E.Network):
    def __init__(self, name, filters=64, cin=1, dropout_rate=0.2, noise_std=0.1, seed=None):
        super().__init__(name, disc_param_shapes(filters, cin))
        f = filters
        self.noise_std, self.rate = noise_std, dropout_rate
        self.conv0 = E.Conv3D(self, "d0.conv", 4, 2, cin, f)
        self.norm0 = E.InstanceNorm(self, "d0.in", f)
        self.conv1 = E.Conv3D(self, "d1.conv", 4, 2, f, 2 * f, use_bias=False)
        self.norm1 = E.InstanceNorm(self, "d1.in", 2 * f)
        self.conv2 = E.Conv3D(self, "d2.conv", 4, 2, 2 * f, 4 * f, use_bias=False)
        self.norm2 = E.InstanceNorm(self, "d2.in", 4 * f)
        self.conv3 = E.Conv3D(self, "d3.conv", 4, 1, 4 * f, 8 * f, use_bias=False, dx_crop=(1, 2))   # TF 'same' zeros
        self.norm3 = E.InstanceNorm(self, "d3.in", 8 * f)
        self.convo = E.Conv3D(self, "dout.conv", 3, 1, 8 * f, 1)
        self.rng_step = 0
        if seed is not None:
            self.load(E.default_init({n: p.shape for n, p in self.params.items()}, seed, glorot=tuple(n for n in self.params if n.endswith(".in.gamma"))))

    def stage(self, k, tape, h, training=True, noise=None, masks=None, seed=0, seed_dev=None):
        """Stage k = 0..4: everything between the raw output of convolution k-1 (the network input for k = 0) and the raw
        output of convolution k, i.e. InstanceNorm + LeakyReLU + SpatialDropout3D of block k-1, then padding + GaussianNoise and
        the convolution of block k (discriminator.py:50-114).  `forward` chains the five stages; the teacher-forced parity
        tests run them one at a time."""
        n = h.shape[0]
        std = self.noise_std if training else 0.0
        nz = None if noise is None else noise[k]

        def mask(i, c):
            if not training:
                return None
            if masks is not None:
                return masks[i].reshape(n * c).contiguous()
            return E.dropout_mask(n * c, self.rate, seed * 16 + 8 + i, seed_dev)

        kw = dict(noise=nz, noise_std=std, seed=seed * 16 + k, seed_dev=seed_dev)
        if k == 0:
            return self.conv0(tape, E.pad_noise(tape, h, **kw))
        if k == 1:
            return self.conv1(tape, self.norm0(tape, h, act=ACT_LEAKY, pad=(1, 1, PAD_REFLECT), **kw))
        if k == 2:
            return self.conv2(tape, self.norm1(tape, h, act=ACT_LEAKY, pad=(1, 1, PAD_REFLECT), drop=mask(0, self.norm1.c), **kw))
        if k == 3:
            # next conv is k4 s1 'same': TF pads 1 before / 2 after with zeros, AFTER the noise layer
            return self.conv3(tape, self.norm2(tape, h, act=ACT_LEAKY, pad=(1, 2, PAD_ZERO), drop=mask(1, self.norm2.c), **kw))
        return self.convo(tape, self.norm3(tape, h, act=ACT_LEAKY, pad=(1, 1, PAD_ZERO), drop=mask(2, self.norm3.c), **kw))

    def forward(self, tape, x, training=True, noise=None, masks=None, seed=0, seed_dev=None):
        """x: Var (N,D,H,W,1) fp32.  noise / masks: explicit tensors (parity mode; oracle layout) or None
        -> in-kernel Philox noise and channel masks keyed on `seed` (+ the per-step offset *seed_dev, a device scalar)."""
        h = x
        for k in range(5):
            h = self.stage(k, tape, h, training=training, noise=noise, masks=masks, seed=seed, seed_dev=seed_dev)
        return h

    def __call__(self, x, training=False):
        xt = torch.as_tensor(x, dtype=torch.float32, device=E.DEV).contiguous()
        return self.forward(E.Tape(enabled=False), E.Var(xt), training=training, seed=self.rng_step).data


def get_discriminator(input_img_size=(64, 64, 512, 1), batch_size=None, filters=64, kernel_initializer='he_normal',
                      num_downsampling=3, use_dropout=False, dropout_rate=0.2, wasserstein=False, use_SN=False,
                      use_input_noise=False, use_layer_noise=False, use_standardisation=False, name=None,
                      noise_std=0.1, seed=0):
    """Same arguments as discriminator.py:7-22.  Built for the reference's live configuration
    (use_dropout, use_input_noise, use_layer_noise all True; no SN; LSGAN head)."""
    if wasserstein or use_SN or num_downsampling != 3:
        raise NotImplementedError("discriminator: SpectralNormalization / Wasserstein head / depth != 3 not built")
    if not (use_dropout and use_input_noise and use_layer_noise):
        raise NotImplementedError("discriminator: built for use_dropout/use_input_noise/use_layer_noise=True (vangan.py:167-192)")
    return DiscriminatorModel(name or "discriminator", filters=filters, cin=input_img_size[-1], dropout_rate=dropout_rate,
                              noise_std=noise_std, seed=seed)
