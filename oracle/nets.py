"""Oracle: ResUNet generator + 3D PatchGAN discriminator (torch CPU, autograd).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Tensors are NDHWC at the
interface, exactly like the reference's Keras `channels_last` models; weights are
kept in Keras layout: Conv3D kernel (kd,kh,kw,Cin,Cout), bias (Cout), InstanceNorm
gamma/beta (C).

Follows
  resunet_model.py:23-39   norm_act           resunet_model.py:42-66   conv_block
  resunet_model.py:69-100  stem               resunet_model.py:103-143 residual_block
  resunet_model.py:146-182 upsample_concat    resunet_model.py:185-249 ResUNet
  discriminator.py:47-124  get_discriminator  building_blocks.py:126-196 downsample
  building_blocks.py:15-39 ReflectionPadding3D
with the arguments VanGan passes (vangan.py:112-122,151-162,167-192).
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

IN_EPS = 1e-3  # tfa.layers.InstanceNormalization default epsilon


# --------------------------------------------------------------------------- helpers
def _ncdhw(x):
    return x.permute(0, 4, 1, 2, 3)


def _ndhwc(x):
    return x.permute(0, 2, 3, 4, 1)


def reflect_pad(x, p=1):
    """building_blocks.py:28-39 — tf.pad(..., mode='REFLECT') on the three spatial axes."""
    return _ndhwc(F.pad(_ncdhw(x), (p, p, p, p, p, p), mode="reflect"))


def conv3d(x, w, b=None, stride=1, padding="valid"):
    """Keras Conv3D on NDHWC input with kernel (kd,kh,kw,Cin,Cout).

    'same' follows TensorFlow: total = max((ceil(n/s)-1)*s + k - n, 0), before = total//2,
    after = total - before (the extra voxel goes AFTER).
    """
    k = w.shape[0]
    if w.shape[3] > 1:          # Cin == 1 layers run in fp32 on CUDA cores
        w = _qw(w)
    xt = _ncdhw(x)
    if padding == "same":
        pads = []
        for n in x.shape[1:4][::-1]:  # F.pad wants last axis first
            out = -(-n // stride)
            total = max((out - 1) * stride + k - n, 0)
            pads += [total // 2, total - total // 2]
        xt = F.pad(xt, pads)
    wt = w.permute(4, 3, 0, 1, 2)  # -> (Cout, Cin, kd, kh, kw)
    y = _ndhwc(F.conv3d(xt, wt, b, stride=stride))
    return _qa(y) if w.shape[4] > 1 else y      # single-channel outputs (head, logits) stay fp32


def instance_norm(x, gamma, beta):
    """tfa InstanceNormalization (GroupNorm with groups=C): biased variance over D,H,W per (n,c)."""
    mu = x.mean(dim=(1, 2, 3), keepdim=True)
    var = ((x - mu) ** 2).mean(dim=(1, 2, 3), keepdim=True)
    return (x - mu) * torch.rsqrt(var + IN_EPS) * gamma + beta


def upsample2(x):
    """UpSampling3D(size=2): nearest repeat along the three spatial axes."""
    return x.repeat_interleave(2, 1).repeat_interleave(2, 2).repeat_interleave(2, 3)


# --------------------------------------------------------------------------- bf16 storage emulation
class _RoundBoth(torch.autograd.Function):
    """Rounds a tensor to bf16 in the forward pass AND its gradient in the backward pass: models an
    activation that the CUDA path stores in bf16 (its gradient tensor is stored in bf16 too)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


class _RoundFwd(torch.autograd.Function):
    """bf16 operand copy of an fp32 master weight (straight-through gradient)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


class Emu:
    """Switch for the optional bf16-storage emulation.  With `Emu.on = True` the oracle rounds at
    exactly the points where the CUDA path stores bf16 (conv outputs, normalised+padded activations,
    tensor-core weight operands, and the matching gradient tensors).  Used by the graph-level parity
    test: it separates "is the wiring of the four backward sweeps right" from "how much does bf16
    cost against the fp32 reference"."""
    on = False


def _qa(t):
    return _RoundBoth.apply(t) if Emu.on else t


def _qw(w):
    return _RoundFwd.apply(w) if Emu.on else w


# --------------------------------------------------------------------------- parameters
def he_normal(rng, shape):
    """Keras 'he_normal': truncated normal (|z|<=2), stddev = sqrt(2/fan_in)/0.87962566."""
    fan_in = int(np.prod(shape[:-1]))
    std = math.sqrt(2.0 / fan_in) / 0.87962566103423978
    z = rng.standard_normal(size=shape)
    bad = np.abs(z) > 2
    while bad.any():
        z[bad] = rng.standard_normal(size=int(bad.sum()))
        bad = np.abs(z) > 2
    return (z * std).astype(np.float32)


def resunet_param_shapes(filters=16, num_layers=4, cin=1):
    """Ordered {name: shape}.  Naming: '<block>.<layer>.{w,b,gamma,beta}'."""
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


def disc_param_shapes(filters=64, cin=1):
    P = OrderedDict()
    P["d0.conv.w"] = (4, 4, 4, cin, filters); P["d0.conv.b"] = (filters,)
    P["d0.in.gamma"] = (filters,); P["d0.in.beta"] = (filters,)
    c = filters
    for i in (1, 2, 3):
        P["d%d.conv.w" % i] = (4, 4, 4, c, 2 * c)  # use_bias=False (building_blocks.py:136)
        P["d%d.in.gamma" % i] = (2 * c,); P["d%d.in.beta" % i] = (2 * c,)
        c *= 2
    P["dout.conv.w"] = (3, 3, 3, c, 1); P["dout.conv.b"] = (1,)
    return P


def init_params(shapes, seed, perturb=0.0):
    """He-normal kernels, zero bias, gamma=1, beta=0 (Keras defaults).  `perturb` adds N(0,perturb)
    to bias/gamma/beta so parity tests exercise them (all-zero biases hide bugs)."""
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for name, shp in shapes.items():
        if name.endswith(".w"):
            out[name] = he_normal(rng, shp)
        elif name.endswith(".gamma"):
            out[name] = (1.0 + perturb * rng.standard_normal(shp)).astype(np.float32)
        else:
            out[name] = (perturb * rng.standard_normal(shp)).astype(np.float32)
    return out


def to_torch(params, dtype=torch.float32, requires_grad=True):
    return OrderedDict((k, torch.tensor(v, dtype=dtype, requires_grad=requires_grad)) for k, v in params.items())


# --------------------------------------------------------------------------- ResUNet
def _norm_act(p, name, x, act=True):
    y = instance_norm(x, p[name + ".gamma"], p[name + ".beta"])
    return torch.relu(y) if act else y


def _conv_block(p, name, x, stride=1):
    y = _qa(_norm_act(p, name + ".in", x))
    y = reflect_pad(y)
    return conv3d(y, p[name + ".conv.w"], p[name + ".conv.b"], stride=stride)


def _res_block(p, name, x, stride):
    res = _conv_block(p, name + ".cb1", x, stride)
    res = _conv_block(p, name + ".cb2", res, 1)
    sc = conv3d(x, p[name + ".short.conv.w"], p[name + ".short.conv.b"], stride=stride, padding="same")
    sc = _norm_act(p, name + ".short.in", sc, act=False)
    return _qa(sc + res)


def resunet_forward(p, x, num_layers=4, taps=None):
    """ResUNet(upsample_mode='simple', dropout_type='none', filters=16, num_layers=4, tanh head).
    `taps`: optional dict that receives named intermediate tensors (for layer-wise parity tests)."""
    conv = conv3d(reflect_pad(x), p["stem.conv0.w"], p["stem.conv0.b"])
    conv = _conv_block(p, "stem.cb", conv)
    sc = conv3d(x, p["stem.short.conv.w"], p["stem.short.conv.b"], padding="same")
    sc = _norm_act(p, "stem.short.in", sc, act=False)
    h = _qa(conv + sc)
    skips = [h]
    if taps is not None:
        taps["stem"] = h
    for e in range(1, num_layers + 1):
        h = _res_block(p, "enc%d" % e, h, 2)
        skips.append(h)
        if taps is not None:
            taps["enc%d" % e] = h
    h = _conv_block(p, "bridge1", h)
    h = _conv_block(p, "bridge2", h)
    if taps is not None:
        taps["bridge"] = h
    for d in reversed(range(num_layers)):
        h = torch.cat([upsample2(h), skips[d]], dim=-1)
        h = _res_block(p, "dec%d" % d, h, 1)
        if taps is not None:
            taps["dec%d" % d] = h
    return torch.tanh(conv3d(h, p["head.w"], p["head.b"], padding="same"))


# --------------------------------------------------------------------------- discriminator
def disc_noise_shapes(n, s, filters=64):
    """Shapes of the 5 GaussianNoise tensors and 3 SpatialDropout3D masks for an n x s^3 x 1 input."""
    s1, s2, s3 = s // 2, s // 4, s // 8
    noise = [(n, s + 2, s + 2, s + 2, 1), (n, s1 + 2, s1 + 2, s1 + 2, filters),
             (n, s2 + 2, s2 + 2, s2 + 2, 2 * filters), (n, s3, s3, s3, 4 * filters),
             (n, s3, s3, s3, 8 * filters)]
    masks = [(n, 1, 1, 1, 2 * filters), (n, 1, 1, 1, 4 * filters), (n, 1, 1, 1, 8 * filters)]
    return noise, masks


def make_disc_rand(rng, n, s, filters=64, noise_std=0.1, rate=0.2, dtype=torch.float32):
    """Explicit noise / dropout tensors (the reference draws them from TF's stateful RNG; parity
    tests inject the same tensors into the oracle and the CUDA path)."""
    ns, ms = disc_noise_shapes(n, s, filters)
    noise = [torch.tensor(noise_std * rng.standard_normal(sh), dtype=dtype) for sh in ns]
    masks = [torch.tensor((rng.random(sh) >= rate) / (1.0 - rate), dtype=dtype) for sh in ms]
    return noise, masks


def disc_forward(p, x, noise=None, masks=None, taps=None):
    """get_discriminator(filters=64, use_dropout=True, dropout_rate=0.2, use_input_noise=True,
    use_layer_noise=True, noise_std=0.1) in training mode when noise/masks are given; inference
    mode (no noise, no dropout) when they are None."""
    def nz(i, t):
        return t if noise is None else t + noise[i]

    def dr(i, t):
        return t if masks is None else t * masks[i]

    h = nz(0, reflect_pad(x))                                         # discriminator.py:50-52
    h = conv3d(h, p["d0.conv.w"], p["d0.conv.b"], stride=2)           # :63-69
    h = F.leaky_relu(instance_norm(h, p["d0.in.gamma"], p["d0.in.beta"]), 0.2)   # :70-72
    if taps is not None:
        taps["d0"] = h
    for i in (1, 2):                                                  # :75-88 -> downsample()
        h = _qa(nz(i, reflect_pad(h)))
        h = conv3d(h, p["d%d.conv.w" % i], None, stride=2)
        h = F.leaky_relu(instance_norm(h, p["d%d.in.gamma" % i], p["d%d.in.beta" % i]), 0.2)
        h = dr(i - 1, h)
        if taps is not None:
            taps["d%d" % i] = h
    h = _qa(nz(3, h))                                                 # :90-103, padding='same'
    h = conv3d(h, p["d3.conv.w"], None, stride=1, padding="same")
    h = F.leaky_relu(instance_norm(h, p["d3.in.gamma"], p["d3.in.beta"]), 0.2)
    h = dr(2, h)
    if taps is not None:
        taps["d3"] = h
    h = _qa(nz(4, h))                                                 # :105-114
    return conv3d(h, p["dout.conv.w"], p["dout.conv.b"], padding="same")


# --------------------------------------------------------------------------- V-Net (both variants VanGan builds)
BN_EPS, BN_MOMENTUM = 1e-3, 0.99   # keras.layers.BatchNormalization defaults (vnet_model.py:127-128 passes none)


def vnet_param_shapes(filters=32, num_layers=4, cin=1, use_batch_norm=False, deconv=False):
    """Trainable variables of custom_vnet in call order (vnet_model.py:199-264).  gen_IS arguments (vangan.py:97-110):
    use_batch_norm=False, upsample_mode='upsample', filters=32.  gen_SI arguments (vangan.py:135-149): use_batch_norm=True (block
    convolutions get use_bias=False, vnet_model.py:124,139), upsample_mode='deconv' (Conv3DTranspose kernel (2,2,2,Cout,Cin) + bias,
    vnet_model.py:245), filters=16."""
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
        if deconv:
            P["dec%d.up.w" % l] = (2, 2, 2, f, 2 * f)
            P["dec%d.up.b" % l] = (f,)
        else:
            conv("dec%d.up.conv" % l, 3, 2 * f, f)
        block("dec%d" % l, 2 * f, f)
    conv("head", 1, f, 1)
    return P


def make_vnet_masks(rng, n, filters=32, num_layers=4, rate=0.5, dtype=torch.float32):
    """SpatialDropout3D(0.5) masks (one per encoder block + bottleneck), scaled by 1/(1-rate) (vnet_model.py:131-132)."""
    return [torch.tensor((rng.random((n, 1, 1, 1, filters * 2 ** l)) >= rate) / (1.0 - rate), dtype=dtype)
            for l in range(num_layers + 1)]


def batch_norm(x, gamma, beta, state=None, key=None, training=True):
    """keras.layers.BatchNormalization(axis=-1, momentum=0.99, epsilon=1e-3) on NDHWC.  training: biased batch statistics over
    N,D,H,W; the moving averages in `state` ({key + '.moving_mean' / '.moving_variance'}) are updated in place with
    moving = m*moving + (1-m)*batch, the variance Bessel-corrected (Keras' fused 5-D path, recalled).  inference: the moving values."""
    if training:
        mu = x.mean(dim=(0, 1, 2, 3), keepdim=True)
        var = ((x - mu) ** 2).mean(dim=(0, 1, 2, 3), keepdim=True)
        if state is not None:
            cnt = x.numel() // x.shape[-1]
            with torch.no_grad():
                mm, mv = state[key + ".moving_mean"], state[key + ".moving_variance"]
                mm.mul_(BN_MOMENTUM).add_((1 - BN_MOMENTUM) * mu.reshape(-1).to(mm.dtype))
                mv.mul_(BN_MOMENTUM).add_((1 - BN_MOMENTUM) * (var.reshape(-1) * (cnt / max(cnt - 1, 1))).to(mv.dtype))
    else:
        mu, var = state[key + ".moving_mean"].to(x.dtype), state[key + ".moving_variance"].to(x.dtype)
    return (x - mu) * torch.rsqrt(var + BN_EPS) * gamma + beta


def conv3d_transpose_k2s2(x, w, b):
    """Keras Conv3DTranspose(filters, (2,2,2), strides 2, 'same') on NDHWC with kernel (2,2,2,Cout,Cin):
    y[n, 2d+a, 2h+b, 2w+c, co] = sum_ci x[n,d,h,w,ci] * w[a,b,c,co,ci] + bias[co]."""
    wt = _qw(w).permute(4, 3, 0, 1, 2)          # -> (Cin, Cout, kd, kh, kw)
    return _qa(_ndhwc(F.conv_transpose3d(_ncdhw(x), wt, b, stride=2)))


def vnet_bn_state(filters=16, num_layers=4, dtype=torch.float32):
    """Fresh moving statistics (zeros / ones) of every BatchNormalization, keyed like the CUDA network's buffers."""
    st = OrderedDict()
    names = ["enc%d" % l for l in range(num_layers)] + ["bridge"] + ["dec%d" % l for l in range(num_layers)]
    widths = [filters * 2 ** l for l in range(num_layers)] + [filters * 2 ** num_layers] + [filters * 2 ** l for l in range(num_layers)]
    for nme, c in zip(names, widths):
        for j in (1, 2):
            st["%s.c%d.bn.moving_mean" % (nme, j)] = torch.zeros(c, dtype=dtype)
            st["%s.c%d.bn.moving_variance" % (nme, j)] = torch.ones(c, dtype=dtype)
    return st


def _vnet_block(p, name, x, mask=None, bn_state=None, training=True):
    """conv3d_block (vnet_model.py:80-146): pad -> Conv3D(relu) -> norm -> [SpatialDropout3D] -> pad -> Conv3D(relu) -> norm, norm =
    InstanceNormalization or (bn_state given) BatchNormalization."""
    bn = (name + ".c1.bn.gamma") in p

    def norm(c, j):
        if bn:
            k = "%s.c%d.bn" % (name, j)
            return batch_norm(c, p[k + ".gamma"], p[k + ".beta"], bn_state, k, training)
        k = "%s.c%d.in" % (name, j)
        return instance_norm(c, p[k + ".gamma"], p[k + ".beta"])

    c = torch.relu(conv3d(reflect_pad(x), p[name + ".c1.conv.w"], p.get(name + ".c1.conv.b")))
    c = norm(c, 1)
    if mask is not None:
        c = c * mask
    c = _qa(c)
    c = torch.relu(conv3d(reflect_pad(c), p[name + ".c2.conv.w"], p.get(name + ".c2.conv.b")))
    return _qa(norm(c, 2))


def vnet_forward(p, x, num_layers=4, masks=None, taps=None, bn_state=None, training=None):
    """custom_vnet(..., dropout=0.5, num_layers=4, output_activation='tanh') (vnet_model.py:199-264); the variant (InstanceNorm +
    UpSampling3D/Conv3D, or BatchNorm + Conv3DTranspose) follows from the variables present in `p`.  masks=None is inference mode for
    dropout; `training` (default: masks is not None) selects batch vs moving statistics of BatchNormalization."""
    training = (masks is not None) if training is None else training
    down = []
    for l in range(num_layers):
        x = _vnet_block(p, "enc%d" % l, x, None if masks is None else masks[l], bn_state, training)
        down.append(x)
        if taps is not None:
            taps["enc%d" % l] = x
        x = _ndhwc(F.max_pool3d(_ncdhw(x), 2))                      # MaxPooling3D((2,2,2))
    x = _vnet_block(p, "bridge", x, None if masks is None else masks[num_layers], bn_state, training)
    if taps is not None:
        taps["bridge"] = x
    for l in reversed(range(num_layers)):
        if ("dec%d.up.w" % l) in p:
            x = conv3d_transpose_k2s2(x, p["dec%d.up.w" % l], p["dec%d.up.b" % l])
        else:
            x = conv3d(upsample2(x), p["dec%d.up.conv.w" % l], p["dec%d.up.conv.b" % l], padding="same")
        x = torch.cat([x, down[l]], dim=-1)
        x = _vnet_block(p, "dec%d" % l, x, None, bn_state, training)
        if taps is not None:
            taps["dec%d" % l] = x
    return torch.tanh(conv3d(x, p["head.w"], p["head.b"], padding="same"))


# --------------------------------------------------------------------------- 3D ResNet generator ('resnet', VanGan's default)
def resnet_param_shapes(filters=32, num_downsampling_blocks=3, num_residual_blocks=6, num_upsample_blocks=3, cin=1):
    """Trainable variables of get_resnet_generator (generator.py:7-73) in call order, with VanGan's arguments (vangan.py:88-95: three
    downsampling and three upsampling blocks).  No convolution but the last has a bias (use_bias=False: generator.py:38,
    building_blocks.py:77,135,248)."""
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


def make_resnet_masks(rng, n, filters=32, num_downsampling_blocks=3, dtype=torch.float32):
    """SpatialDropout3D masks in call order: rate 0.5 after the first block (generator.py:42), 0.2 inside every downsample block
    (building_blocks.py:192-195), scaled by 1/(1-rate)."""
    rates = [0.5] + [0.2] * num_downsampling_blocks
    return [torch.tensor((rng.random((n, 1, 1, 1, filters * 2 ** i)) >= r) / (1.0 - r), dtype=dtype) for i, r in enumerate(rates)]


def resnet_stage_c0(p, x, mask=None):
    """generator.py:35-42: ReflectionPadding3D(1) -> Conv3D(7, valid) -> InstanceNorm -> ReLU -> SpatialDropout3D(0.5)"""
    h = torch.relu(instance_norm(conv3d(reflect_pad(x), p["c0.conv.w"]), p["c0.in.gamma"], p["c0.in.beta"]))
    return _qa(h if mask is None else h * mask)


def resnet_stage_down(p, i, x, mask=None):
    """building_blocks.downsample as called at generator.py:45-50: pad -> Conv3D(k3, s2, valid) -> IN -> ReLU -> SpatialDropout3D(0.2)"""
    h = conv3d(reflect_pad(x), p["down%d.conv.w" % i], stride=2)
    h = torch.relu(instance_norm(h, p["down%d.in.gamma" % i], p["down%d.in.beta" % i]))
    return _qa(h if mask is None else h * mask)


def resnet_stage_res(p, j, x):
    """building_blocks.residual_block (:68-123)"""
    h = conv3d(reflect_pad(x), p["res%d.c1.conv.w" % j])
    h = _qa(torch.relu(instance_norm(h, p["res%d.c1.in.gamma" % j], p["res%d.c1.in.beta" % j])))
    h = conv3d(reflect_pad(h), p["res%d.c2.conv.w" % j])
    return _qa(x + instance_norm(h, p["res%d.c2.in.gamma" % j], p["res%d.c2.in.beta" % j]))


def resnet_stage_up(p, i, x):
    """building_blocks.upsample (:240-280): UpSampling3D(2) -> Conv3D(k4, s1, 'same') -> IN -> ReLU"""
    h = conv3d(_qa(upsample2(x)), p["up%d.conv.w" % i], padding="same")
    return _qa(torch.relu(instance_norm(h, p["up%d.in.gamma" % i], p["up%d.in.beta" % i])))


def resnet_stage_out(p, x):
    """generator.py:66-69 (num_downsampling_blocks == 3: no extra padding): Conv3D(1, 7, 'same') -> tanh"""
    return torch.tanh(conv3d(x, p["out.conv.w"], p["out.conv.b"], padding="same"))


def resnet_forward(p, x, nd=3, nr=6, nu=3, masks=None, taps=None):
    """get_resnet_generator(num_downsampling_blocks=3, num_upsample_blocks=3) (generator.py:7-73).  masks=None: inference."""
    h = resnet_stage_c0(p, x, None if masks is None else masks[0])
    if taps is not None:
        taps["c0"] = h
    for i in range(nd):
        h = resnet_stage_down(p, i, h, None if masks is None else masks[i + 1])
        if taps is not None:
            taps["down%d" % i] = h
    for j in range(nr):
        h = resnet_stage_res(p, j, h)
        if taps is not None:
            taps["res%d" % j] = h
    for i in range(nu):
        h = resnet_stage_up(p, i, h)
        if taps is not None:
            taps["up%d" % i] = h
    return resnet_stage_out(p, h)
