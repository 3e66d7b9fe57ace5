"""Python side of the TensorFlow custom-op shim (integration/vangan_tf_ops.cc): gradient registrations and the Keras layer
a maintainer of the reference swaps in for `layers.Conv3D` (resunet_model.py:64, discriminator.py:63,
building_blocks.py:182) and `tfa.layers.InstanceNormalization` + activation + padding (resunet_model.py:23-39).

NOT RUN IN THIS IMAGE (TensorFlow is not installed here); it documents the binding and is kept next to the C++ source so
that both follow include/vangan_b200.h.  Usage in the reference:

    from integration.vangan_tf import VgConv3D, VgInstanceNorm, vg_soft_skel
    x = VgInstanceNorm(act=1, pad=(1, 1, 1))(x)          # InstanceNormalization -> ReLU -> ReflectionPadding3D
    x = VgConv3D(16, 3)(x)                               # Conv3D(16, (3,3,3), padding='valid') on the padded tensor
"""
import tensorflow as tf

_ops = tf.load_op_library("vangan_tf_ops.so")


class VgConv3D(tf.keras.layers.Layer):
    """keras.layers.Conv3D(filters, k, strides, padding='valid') over an explicitly padded NDHWC tensor.  Weights keep the Keras
    layout (kd,kh,kw,Cin,Cout) + (Cout,), so reference checkpoints load unchanged."""

    def __init__(self, filters, k, strides=1, act=0, dx_crop=(0, 0), use_bias=True, **kw):
        super().__init__(**kw)
        self.filters, self.k, self.strides, self.act, self.dx_crop, self.use_bias = filters, k, strides, act, dx_crop, use_bias

    def build(self, shape):
        cin = int(shape[-1])
        self.kernel = self.add_weight("kernel", (self.k,) * 3 + (cin, self.filters), initializer="he_normal")
        self.bias = self.add_weight("bias", (self.filters,), initializer="zeros", trainable=self.use_bias)
        self.attrs = dict(k=self.k, stride=self.strides, cout=self.filters, act=self.act, dx_lo=self.dx_crop[0], dx_hi=self.dx_crop[1])

    def call(self, x):
        attrs, cin = self.attrs, int(x.shape[-1])

        @tf.custom_gradient
        def conv(x, kernel, bias):
            w_fwd, w_dgrad = _ops.vg_conv3d_pack(kernel, cin=cin, **attrs)      # cache per optimizer step in production
            y = _ops.vg_conv3d_fwd(x, kernel if cin == 1 else w_fwd, bias, **attrs)

            def grad(dy):
                dx = _ops.vg_conv3d_dgrad(dy, w_dgrad, tf.shape(x), **attrs)
                dw, db = _ops.vg_conv3d_wgrad(x, dy, **attrs)
                return dx, dw, db
            return y, grad
        return conv(x, self.kernel, self.bias)


class VgInstanceNorm(tf.keras.layers.Layer):
    """tfa.layers.InstanceNormalization (eps 1e-3) fused with the activation, residual Add, SpatialDropout3D mask and the
    padding of the next convolution.  act: 0 none / 1 ReLU / 2 LeakyReLU(slope); pad = (lo, hi, mode) with mode 1 = REFLECT."""

    def __init__(self, act=0, slope=0.2, pad=(0, 0, 0), **kw):
        super().__init__(**kw)
        self.attrs = dict(act=act, slope=slope, pad_lo=pad[0], pad_hi=pad[1], pad_mode=pad[2], noise_std=0.0)

    def build(self, shape):
        c = int(shape[-1])
        self.gamma = self.add_weight("gamma", (c,), initializer="ones")
        self.beta = self.add_weight("beta", (c,), initializer="zeros")

    def call(self, x, residual=None, drop=None):
        attrs = self.attrs
        res = residual if residual is not None else tf.zeros((1,), x.dtype)
        drp = drop if drop is not None else tf.zeros((0,), tf.float32)

        @tf.custom_gradient
        def norm(x, res, gamma, beta):
            y, mean, rstd = _ops.vg_instance_norm(x, res, gamma, beta, drp, tf.constant(0, tf.int64), **attrs)

            def grad(dy):
                dx, dres, dg, db = _ops.vg_instance_norm_grad(dy, x, mean, rstd, gamma, beta, drp, **attrs)
                return dx, (dres if residual is not None else tf.zeros_like(res)), dg, db
            return y, grad
        return norm(x, res, self.gamma, self.beta)


@tf.custom_gradient
def _soft_skel(img, iters):
    skel, e, s = _ops.vg_soft_skel(img, iters=iters)
    return skel, lambda g: (_ops.vg_soft_skel_grad(g, e, s, iters=iters), None)


def vg_soft_skel(img, iters):
    """clDice_func.soft_skel(img, iters) (clDice_func.py:60-80) on the fused kernels."""
    return _soft_skel(img, iters)
