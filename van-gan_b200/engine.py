"""Host-side execution engine: a persistent gradient tape over the C-ABI kernels.

Plays the role tf.GradientTape(persistent=True) + Keras layers play in the reference
(vangan.py:394-438): every op records a node; `Tape.backward` runs one reverse sweep for ONE loss
w.r.t. ONE network's variables (what each `optimizer.minimize(loss, var_list, tape)` call does),
visiting only nodes that lie between the seeds and the requested variables.

All device arithmetic happens inside libvangan_b200.so; torch supplies allocation and streams.
Activations are NDHWC, bf16 for multi-channel feature maps, fp32 for single-channel volumes.
"""
import math

import numpy as np
import torch

from . import _lib
from ._lib import ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_TANH, PAD_REFLECT, PAD_ZERO, ConvDesc, InDesc, call, dtype_code

DEV = "cuda"
BIAS_SINKS = __import__("os").environ.get("VG_BIAS_SINKS", "1") != "0"   # conv bias gradients taken inside the consuming norm's backward
STATS_REUSE = __import__("os").environ.get("VG_STATS_REUSE", "1") != "0"   # InstanceNorm statistics cached per tensor / derived for concats


# ----------------------------------------------------------------------------- tape
class Var:
    """A tensor on the tape.  `data`: torch CUDA tensor; `grad`: accumulated gradient or None."""
    __slots__ = ("data", "grad", "src", "ncons", "bias_sink", "bias_done", "stats")

    def __init__(self, data, src=None):
        self.data = data
        self.grad = None
        self.src = src
        self.ncons = 0           # number of recorded consumers
        self.bias_sink = None    # bias Param of the convolution that produced this tensor (its gradient is the channel sum of ours)
        self.bias_done = False   # that channel sum was already taken by the consumer's backward kernel in this sweep
        self.stats = None        # (mean, rstd) per (n, c) of this tensor once an InstanceNorm has computed them (re-used, see upsample_concat)

    @property
    def shape(self):
        return tuple(self.data.shape)


class Node:
    __slots__ = ("inputs", "outputs", "params", "bwd", "name")

    def __init__(self, inputs, outputs, params, bwd, name):
        self.inputs, self.outputs, self.params, self.bwd, self.name = inputs, outputs, params, bwd, name


def accumulate(var, g):
    """var.grad += g (takes ownership of g when it is the first contribution)."""
    if var.grad is None:
        var.grad = g
    else:
        call("vg_accumulate", var.grad, g, g.numel(), dtype_code(g))


class Tape:
    KEEP_BYTES = 2 << 30     # weight-gradient stream: join once this many bytes of incoming gradients are being held for it

    def __init__(self, enabled=True):
        self.nodes = []
        self.enabled = enabled
        self.wrt = set()          # ids of the Params the running sweep differentiates
        self.wg_stream = None     # optional sibling stream of the current backward sweep for the weight-gradient kernels
        self._keep, self._keep_bytes = [], 0

    # A convolution's weight gradient and its input gradient both consume dy and nothing else of each other, so the weight-gradient
    # kernel can run on a sibling stream beside the rest of the sweep.  dy is a tape temporary: it is kept referenced until the
    # sweep stream has been ordered after the sibling again (only then may the allocator hand its memory to later kernels).
    def side_call(self, hold, fn):
        if self.wg_stream is None:
            fn()
            return
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        self.wg_stream.wait_event(ev)
        with torch.cuda.stream(self.wg_stream):
            fn()
        self._keep.append(hold)
        self._keep_bytes += sum(t.numel() * t.element_size() for t in hold)
        if self._keep_bytes > self.KEEP_BYTES:
            self.side_join()

    def side_join(self):
        if self.wg_stream is not None and self._keep:
            ev = torch.cuda.Event()
            ev.record(self.wg_stream)
            torch.cuda.current_stream().wait_event(ev)
        self._keep, self._keep_bytes = [], 0

    def record(self, inputs, outputs, params, bwd, name):
        if self.enabled:
            node = Node(inputs, outputs, params, bwd, name)
            for o in outputs:
                o.src = node
            for v in inputs:
                v.ncons += 1
            self.nodes.append(node)

    def clear(self):
        """Break the Var <-> Node reference cycles so the step's activations are released immediately
        (by reference counting, without waiting for Python's cyclic GC)."""
        for node in self.nodes:
            for o in node.outputs:
                o.src = None
                o.grad = None
            node.bwd = None
            node.inputs = node.outputs = node.params = ()
        self.nodes = []

    def backward(self, seeds, wrt_params, wrt_vars=()):
        """seeds: list of (Var, grad tensor).  wrt_params: Param objects whose .grad views receive
        (+=) the gradients; wrt_vars: leaf Vars whose .grad is wanted as well (kept after the sweep).
        One reverse sweep; gradients of intermediate Vars are released as soon as they are consumed."""
        wrt = set(id(p) for p in wrt_params)
        self.wrt = wrt
        needs = set(id(v) for v in wrt_vars)
        keep = set(needs)
        for node in self.nodes:
            if any(id(p) in wrt for p in node.params) or any(id(v) in needs for v in node.inputs):
                for o in node.outputs:
                    needs.add(id(o))
        for v, g in seeds:
            accumulate(v, g)
        for node in reversed(self.nodes):
            outs = node.outputs
            if all(o.grad is None for o in outs):
                continue
            if id(outs[0]) not in needs:
                for o in outs:
                    o.grad = None
                continue
            in_needs = [id(v) in needs for v in node.inputs]
            p_needs = any(id(p) in wrt for p in node.params)
            node.bwd(in_needs, p_needs)
            for o in outs:
                o.grad = None
        self.side_join()
        # inputs that are leaves keep their grads; clear everything else that may linger
        for node in self.nodes:
            for v in node.inputs:
                if v.src is None and id(v) not in keep:
                    v.grad = None


# ----------------------------------------------------------------------------- parameters
class Param:
    """A trainable variable: fp32 views into its network's flat weight / gradient buffers."""
    __slots__ = ("name", "shape", "offset", "size", "w", "grad")

    def __init__(self, name, shape, offset):
        self.name, self.shape, self.offset = name, tuple(shape), offset
        self.size = int(np.prod(shape))
        self.w = None
        self.grad = None


class Network:
    """Owns the flat fp32 master weights, gradients and Adam slots of one Keras model, in the
    model's trainable_variables order (Keras layouts, so a TF checkpoint maps 1:1)."""

    def __init__(self, name, shapes):
        self.name = name
        self.params = {}
        off = 0
        for n, shp in shapes.items():
            self.params[n] = Param(n, shp, off)
            off += self.params[n].size
        self.total = off
        self.w = torch.zeros(off, dtype=torch.float32, device=DEV)
        self.g = torch.zeros(off, dtype=torch.float32, device=DEV)
        self.m = torch.zeros(off, dtype=torch.float32, device=DEV)
        self.v = torch.zeros(off, dtype=torch.float32, device=DEV)
        offs = [p.offset for p in self.params.values()] + [off]
        self.seg = torch.tensor(offs, dtype=torch.int64, device=DEV)
        self.norm_ws = torch.zeros(len(self.params), dtype=torch.float64, device=DEV)
        for p in self.params.values():
            p.w = self.w[p.offset:p.offset + p.size].view(p.shape)
            p.grad = self.g[p.offset:p.offset + p.size].view(p.shape)
        self.convs = []
        self._pack = False        # operand-pack job table: built lazily, invalidated when a layer registers
        self.pre_pack = []        # callables run before the operand pack (Conv3DTranspose: Keras layout -> GEMM layout)
        self.buffers = {}         # non-trainable state by name (BatchNormalization moving_mean / moving_variance), checkpointed
        self.step_count = 0

    @property
    def trainable_variables(self):
        return list(self.params.values())

    def load(self, arrays):
        """arrays: {name: numpy array in Keras layout}."""
        for n, p in self.params.items():
            a = np.asarray(arrays[n], dtype=np.float32)
            assert a.shape == p.shape, (n, a.shape, p.shape)
            p.w.copy_(torch.from_numpy(a).to(DEV))
        self.repack()

    def export(self):
        return {n: p.w.detach().cpu().numpy().copy() for n, p in self.params.items()}

    def export_grads(self):
        return {n: p.grad.detach().cpu().numpy().copy() for n, p in self.params.items()}

    def build_pack_plan(self):
        """Job table of every bf16 operand copy of this network (one entry per packed buffer) + exclusive prefix of their sizes,
        uploaded once; `repack` then refreshes all of them with ONE launch (vg_pack_run)."""
        import ctypes as C
        L = _lib.lib()
        jb = L.vg_pack_job_bytes()
        chunks, totals = [], []
        for c in self.convs:
            if c.wf is None and c.wd is None:
                continue
            buf = C.create_string_buffer(jb * 24)
            n = L.vg_conv3d_pack_jobs(C.byref(c.desc(1, 8, 8, 8)), c.w.w.data_ptr(), c.wf.data_ptr() if c.wf is not None else None,
                                      c.wd.data_ptr() if c.wd is not None else None, buf, 24)
            if n < 0:
                raise _lib.VgError("vg_conv3d_pack_jobs failed (%d)" % n)
            for i in range(n):
                totals.append(L.vg_pack_job_total(buf, i))
            chunks.append(buf.raw[:jb * n])
        if not totals:
            self._pack = None
            return
        prefix = np.concatenate([[0], np.cumsum(totals)]).astype(np.int64)
        raw = np.frombuffer(b"".join(chunks), dtype=np.uint8).copy()
        self._pack = (torch.from_numpy(raw).to(DEV), torch.from_numpy(prefix[:-1].copy()).to(DEV), len(totals), int(prefix[-1]))

    def repack(self):
        """bf16 operand copies <- fp32 master weights (after load / every optimizer step)."""
        for hook in self.pre_pack:
            hook()
        if self._pack is False:
            self.build_pack_plan()
        if self._pack is not None:
            jobs, prefix, n, total = self._pack
            call("vg_pack_run", jobs, prefix, n, total)

    def zero_grad(self):
        self.g.zero_()

    @staticmethod
    def lr_t(t, lr=2e-4, beta_1=0.5, beta_2=0.9):
        """Keras OptimizerV2 Adam step size at iteration t: lr*sqrt(1-b2^t)/(1-b1^t)."""
        return lr * math.sqrt(1.0 - beta_2 ** t) / (1.0 - beta_1 ** t)

    def adam_step(self, lr=2e-4, beta_1=0.5, beta_2=0.9, eps=1e-7, clipnorm=100.0, lr_t_dev=None):
        """Keras OptimizerV2 Adam: per-variable clip_by_norm, lr_t = lr*sqrt(1-b2^t)/(1-b1^t).
        lr_t_dev: optional 1-element device tensor already holding lr_t for THIS step (graph-replay mode: the caller owns
        the step counter and uploads the value before every replay)."""
        if lr_t_dev is not None:
            call("vg_clip_adam_step_dev", self.w, self.g, self.m, self.v, self.seg, len(self.params), self.total, lr_t_dev, beta_1,
                 beta_2, eps, clipnorm, self.norm_ws)
        else:
            self.step_count += 1
            call("vg_clip_adam_step", self.w, self.g, self.m, self.v, self.seg, len(self.params), self.total,
                 self.lr_t(self.step_count, lr, beta_1, beta_2), beta_1, beta_2, eps, clipnorm, self.norm_ws)
        self.repack()


class PolynomialDecay:
    """tf.keras.optimizers.schedules.PolynomialDecay (cycle=False), the schedule GanMonitor.set_learning_rate installs
    (custom_callback.py:343-365): lr(step) = (lr0 - end) * (1 - min(step, decay_steps)/decay_steps)**power + end, evaluated at the
    optimizer's own iteration counter -- which is NOT reset when the schedule is installed, exactly as in Keras."""

    def __init__(self, initial_learning_rate, decay_steps, end_learning_rate=0.0001, power=1.0):
        self.initial_learning_rate, self.decay_steps = float(initial_learning_rate), float(decay_steps)
        self.end_learning_rate, self.power = float(end_learning_rate), float(power)

    def __call__(self, step):
        if self.decay_steps <= 0:
            return self.end_learning_rate
        s = min(float(step), self.decay_steps)
        return (self.initial_learning_rate - self.end_learning_rate) * (1.0 - s / self.decay_steps) ** self.power + self.end_learning_rate


class Adam:
    """tf.keras.optimizers.Adam(lr, beta_1, beta_2, clipnorm) bound to one Network (vangan.py:220-235).  `lr` may be a float or a
    schedule (callable of the iteration count), and may be reassigned between steps like `model.gen_I_optimizer.lr = ...` in
    custom_callback.py:343.  `iterations` is the network's update count (Keras `optimizer.iterations`); the moment slots m / v
    live in the network's flat buffers."""

    def __init__(self, net, learning_rate=2e-4, beta_1=0.5, beta_2=0.9, epsilon=1e-7, clipnorm=100.0):
        self.net, self.lr = net, learning_rate
        self.beta_1, self.beta_2, self.epsilon, self.clipnorm = beta_1, beta_2, epsilon, clipnorm

    @property
    def iterations(self):
        return self.net.step_count

    @property
    def learning_rate(self):
        return self.lr

    @learning_rate.setter
    def learning_rate(self, v):
        self.lr = v

    def current_lr(self):
        """Decayed learning rate of the NEXT update (Keras evaluates the schedule at `iterations` before incrementing)."""
        return float(self.lr(self.net.step_count)) if callable(self.lr) else float(self.lr)

    def step_size(self):
        """lr_t of the next update: lr * sqrt(1 - b2^t) / (1 - b1^t), t = iterations + 1."""
        return Network.lr_t(self.net.step_count + 1, self.current_lr(), self.beta_1, self.beta_2)


def he_normal(rng, shape):
    """Keras 'he_normal' (truncated normal, stddev sqrt(2/fan_in)/0.8796...)."""
    fan_in = int(np.prod(shape[:-1]))
    std = math.sqrt(2.0 / fan_in) / 0.87962566103423978
    z = rng.standard_normal(size=shape)
    bad = np.abs(z) > 2
    while bad.any():
        z[bad] = rng.standard_normal(size=int(bad.sum()))
        bad = np.abs(z) > 2
    return (z * std).astype(np.float32)


def glorot_uniform(rng, shape):
    """Keras' default initializer (what a layer built WITHOUT kernel_initializer / with gamma_initializer=None gets):
    uniform(-l, l), l = sqrt(6 / (fan_in + fan_out)); a 1-D variable of length C has fan_in = fan_out = C."""
    if len(shape) == 1:
        fan_in = fan_out = shape[0]
    else:
        rf = int(np.prod(shape[:-2]))
        fan_in, fan_out = rf * shape[-2], rf * shape[-1]
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def default_init(shapes, seed, glorot=(), he_gamma=False):
    """Initial values by the reference's initializers: he_normal kernels, zero biases / betas, unit gammas; the variables named in
    `glorot` (name or name prefix) are the ones the reference builds with Keras' default glorot_uniform -- the ResUNet stem convolutions
    and head (resunet_model.py:90,92,96,245), the PatchGAN's InstanceNormalization gammas (gamma_initializer=None, discriminator.py:70,
    building_blocks.py:190), the V-Net head and Conv3DTranspose kernels (vnet_model.py:245,264)."""
    rng = np.random.default_rng(seed)
    out = {}
    for n, shp in shapes.items():
        if any(n == g or n.startswith(g) for g in glorot) and (n.endswith(".w") or n.endswith(".gamma")):
            out[n] = glorot_uniform(rng, shp)
        elif n.endswith(".w"):
            out[n] = he_normal(rng, shp)
        elif n.endswith(".gamma"):
            # gamma_initializer='he_normal' (generator.py:14): a 1-D variable has fan_in = its length
            out[n] = he_normal(rng, (shp[0], shp[0]))[0].copy() if he_gamma else np.ones(shp, np.float32)
        else:
            out[n] = np.zeros(shp, np.float32)
    return out


# ----------------------------------------------------------------------------- layers
class Conv3D:
    """keras.layers.Conv3D(filters, k, strides, padding='valid') over an explicitly padded input."""

    def __init__(self, net, name, k, stride, cin, cout, use_bias=True, act=ACT_NONE, dx_crop=(0, 0), w_param=None):
        """dx_crop=(lo, hi): the input carries `lo`/`hi` voxels of ZERO padding per side, so dgrad may skip them.
        w_param: kernel variable to use instead of net.params[name + '.w'] (Conv3DTranspose's GEMM-layout shadow)."""
        self.dx_crop = dx_crop
        self.w = w_param if w_param is not None else net.params[name + ".w"]
        self.b = net.params[name + ".b"] if use_bias else None
        self.k, self.stride, self.cin, self.cout, self.act = k, stride, cin, cout, act
        self.x_dtype = _lib.VG_F32 if cin == 1 else _lib.VG_BF16
        self.y_dtype = _lib.VG_F32 if cout == 1 else _lib.VG_BF16
        d = self.desc(1, 8, 8, 8)
        nb_f = _lib.lib().vg_conv3d_packed_bytes(d, 0)
        nb_d = _lib.lib().vg_conv3d_packed_bytes(d, 1)
        self.wf = torch.empty(max(nb_f, 2) // 2, dtype=torch.bfloat16, device=DEV) if nb_f else None
        self.wd = torch.empty(max(nb_d, 2) // 2, dtype=torch.bfloat16, device=DEV) if nb_d else None
        net.convs.append(self)
        net._pack = False

    def desc(self, n, d, h, w):
        return ConvDesc(n, d, h, w, self.cin, self.cout, self.k, self.stride, self.x_dtype, self.y_dtype, self.act,
                        self.dx_crop[0], self.dx_crop[1])

    def repack(self):
        if self.wf is not None or self.wd is not None:
            call("vg_conv3d_pack_weights", self.desc(1, 8, 8, 8), self.w.w, self.wf, self.wd)

    def __call__(self, tape, x):
        n, d, h, w, c = x.shape
        assert c == self.cin, (c, self.cin)
        desc = self.desc(n, d, h, w)
        od, oh, ow = [(s - self.k) // self.stride + 1 for s in (d, h, w)]
        y = torch.empty((n, od, oh, ow, self.cout), device=DEV,
                        dtype=torch.float32 if self.cout == 1 else torch.bfloat16)
        flops = 2.0 * self.k ** 3 * self.cin * self.cout * n * od * oh * ow
        call("vg_conv3d_fwd", desc, x.data, self.w.w if self.cin == 1 else self.wf, self.b.w if self.b else None, y,
             work=flops)
        out = Var(y)
        if self.b is not None and self.act == ACT_NONE and self.cout > 1:
            out.bias_sink = self.b      # a single consumer that is a norm takes the bias gradient in its own backward pass

        def bwd(in_needs, p_needs):
            dy = out.grad
            if self.act == ACT_TANH:
                t = torch.empty_like(dy)
                call("vg_tanh_bwd", dy, out.data, t, dy.numel())
                dy = t
            if p_needs:
                dbias = self.b.grad if (self.b is not None and not out.bias_done) else None
                tape.side_call((dy,), lambda: call("vg_conv3d_wgrad", desc, x.data, dy, self.w.grad, dbias, work=flops))
            out.bias_done = False
            if in_needs[0]:
                dx = torch.empty_like(x.data)
                call("vg_conv3d_dgrad", desc, dy, self.w.w if self.cout == 1 else self.wd, dx, work=flops)
                accumulate(x, dx)

        tape.record([x], [out], [self.w] + ([self.b] if self.b else []), bwd, "conv")
        return out


class InstanceNorm:
    """tfa.layers.InstanceNormalization followed by the activation / residual Add / SpatialDropout3D /
    GaussianNoise / padding that the reference applies before the next convolution."""

    def __init__(self, net, name, c):
        self.gamma = net.params[name + ".gamma"]
        self.beta = net.params[name + ".beta"]
        self.c = c

    def __call__(self, tape, x, act=ACT_NONE, slope=0.2, residual=None, pad=(0, 0, PAD_ZERO), drop=None, noise=None,
                 noise_std=0.0, seed=0, relu_input=False, seed_dev=None):
        """relu_input: x is the PRE-activation output of a Conv3D(activation='relu') (vnet_model.py:118-126); the ReLU is
        applied on load and its gradient mask on dx, so the convolution kernels never see the activation."""
        n, d, h, w, c = x.shape
        assert c == self.c
        dt = dtype_code(x.data) | (_lib.IN_RELU_INPUT if relu_input else 0)
        ws_bytes = _lib.lib().vg_instnorm_workspace_bytes(n, d, h, w, c)
        ws = torch.empty(ws_bytes // 4 + 1, dtype=torch.float32, device=DEV)
        mean = torch.empty(n * c, dtype=torch.float32, device=DEV)
        rstd = torch.empty(n * c, dtype=torch.float32, device=DEV)
        # algorithmic HBM bytes: forward = read x twice (statistics, apply) + write y; backward = read x and dy twice + write dx
        es, nin = x.data.element_size(), x.data.numel()
        if x.stats is not None and not relu_input and STATS_REUSE:
            mean, rstd = x.stats           # already known: the tensor was normalised before, or is a concat of tensors that were
        else:
            call("vg_instnorm_stats", x.data, dt, n, d, h, w, c, mean, rstd, ws, ws_bytes, work=float(es * nin))
            if not relu_input:
                x.stats = (mean, rstd)
        desc = InDesc(n, d, h, w, c, dt, act, slope, pad[0], pad[1], pad[2], noise_std if noise is None else 0.0, seed,
                      seed_dev.data_ptr() if seed_dev is not None else None)
        pp = pad[0] + pad[1]
        y = torch.empty((n, d + pp, h + pp, w + pp, c), dtype=x.data.dtype, device=DEV)
        call("vg_instnorm_apply", desc, x.data, residual.data if residual is not None else None, y, mean, rstd,
             self.gamma.w, self.beta.w, drop, noise, work=float(es * (nin * (2 if residual is not None else 1) + y.numel())))
        out = Var(y)
        ins = [x] + ([residual] if residual is not None else [])
        # the incoming gradient is a tape temporary that dies with this node: the kernel may fold its reflected halo in place
        desc_b = InDesc(n, d, h, w, c, dt | _lib.IN_DY_SCRATCH, act, slope, pad[0], pad[1], pad[2], noise_std if noise is None else 0.0, seed,
                        seed_dev.data_ptr() if seed_dev is not None else None)

        def bwd(in_needs, p_needs):
            dy = out.grad
            # x already has a gradient from another consumer: add into it in the same pass (no extra accumulate kernel)
            fuse = in_needs[0] and x.grad is not None and x.grad.dtype == x.data.dtype
            dx = x.grad if fuse else torch.empty_like(x.data)
            need_res = residual is not None and in_needs[1]
            dres = torch.empty_like(x.data) if need_res else None
            ws2 = torch.empty(ws_bytes // 4 + 1, dtype=torch.float32, device=DEV)

            def sink(v, ok):
                """bias gradient of the convolution that produced v, when this norm is v's only consumer and the sweep wants it"""
                if ok and v.bias_sink is not None and v.ncons == 1 and id(v.bias_sink) in tape.wrt and BIAS_SINKS:
                    v.bias_done = True
                    return v.bias_sink.grad
                return None
            sink_x = sink(x, in_needs[0] and not fuse)
            sink_r = sink(residual, need_res) if residual is not None else None
            call("vg_instnorm_bwd_sinks", desc_b, dy, x.data, mean, rstd, self.gamma.w, self.beta.w, drop, dx, 1 if fuse else 0, dres,
                 self.gamma.grad if p_needs else None, self.beta.grad if p_needs else None, sink_x, sink_r, ws2, ws_bytes,
                 work=float(es * (2 * (nin + y.numel()) + nin * (2 if need_res else 1))))
            if in_needs[0] and not fuse:
                accumulate(x, dx)
            if need_res:
                accumulate(residual, dres)

        tape.record(ins, [out], [self.gamma, self.beta], bwd, "instnorm")
        return out


class BatchNorm:
    """keras.layers.BatchNormalization() (vnet_model.py:127-128,142-143: axis=-1, momentum=0.99, epsilon=1e-3) followed by the
    SpatialDropout3D / padding the reference applies before the next convolution.  Training: statistics of the LOCAL batch
    (MirroredStrategy keeps BatchNormalization per replica) and the Keras moving-average update; inference: the moving values.
    The arithmetic is InstanceNorm's with the reductions taken over N*D*H*W (vg_batchnorm_stats / vg_batchnorm_bwd)."""

    MOMENTUM = 0.99

    def __init__(self, net, name, c):
        self.gamma = net.params[name + ".gamma"]
        self.beta = net.params[name + ".beta"]
        self.c = c
        self.moving_mean = net.buffers.setdefault(name + ".moving_mean", torch.zeros(c, dtype=torch.float32, device=DEV))
        self.moving_var = net.buffers.setdefault(name + ".moving_variance", torch.ones(c, dtype=torch.float32, device=DEV))

    def __call__(self, tape, x, training=True, act=ACT_NONE, slope=0.2, pad=(0, 0, PAD_ZERO), drop=None, relu_input=False):
        n, d, h, w, c = x.shape
        assert c == self.c
        dt = dtype_code(x.data) | (_lib.IN_RELU_INPUT if relu_input else 0)
        ws_bytes = _lib.lib().vg_batchnorm_workspace_bytes(n, d, h, w, c)
        ws = torch.empty(ws_bytes // 4 + 1, dtype=torch.float32, device=DEV)
        mean = torch.empty(n * c, dtype=torch.float32, device=DEV)
        rstd = torch.empty(n * c, dtype=torch.float32, device=DEV)
        es, nin = x.data.element_size(), x.data.numel()
        call("vg_batchnorm_stats", x.data, dt, n, d, h, w, c, mean, rstd, self.moving_mean, self.moving_var, self.MOMENTUM,
             1 if training else 0, ws, ws_bytes, work=float(es * nin) if training else 0.0)
        desc = InDesc(n, d, h, w, c, dt, act, slope, pad[0], pad[1], pad[2], 0.0, 0, None)
        pp = pad[0] + pad[1]
        y = torch.empty((n, d + pp, h + pp, w + pp, c), dtype=x.data.dtype, device=DEV)
        call("vg_instnorm_apply", desc, x.data, None, y, mean, rstd, self.gamma.w, self.beta.w, drop, None,
             work=float(es * (nin + y.numel())))
        out = Var(y)

        def bwd(in_needs, p_needs):
            assert training, "BatchNormalization backward is only defined for the batch-statistics (training) path"
            dx = torch.empty_like(x.data)
            ws2 = torch.empty(ws_bytes // 4 + 1, dtype=torch.float32, device=DEV)
            call("vg_batchnorm_bwd", desc, out.grad, x.data, mean, rstd, self.gamma.w, self.beta.w, drop, dx, 0, None,
                 self.gamma.grad if p_needs else None, self.beta.grad if p_needs else None, ws2, ws_bytes,
                 work=float(es * (2 * (nin + y.numel()) + nin)))
            if in_needs[0]:
                accumulate(x, dx)

        tape.record([x], [out], [self.gamma, self.beta], bwd, "batchnorm")
        return out


class Conv3DTranspose:
    """keras.layers.Conv3DTranspose(filters, (2,2,2), strides=(2,2,2), padding='same') (vnet_model.py:245).  The eight outputs of an
    input voxel do not overlap, so the layer is ONE pointwise GEMM with 8*Cout columns (a,b,c,co) on the tensor cores (the K = 1
    Conv3D path) followed by a depth-to-space scatter that adds the bias.  The trainable kernel stays in Keras layout
    (2,2,2,Cout,Cin); a GEMM-layout fp32 shadow is refreshed before every operand pack and its gradient is folded back."""

    def __init__(self, net, name, cin, cout):
        self.w = net.params[name + ".w"]           # (2, 2, 2, cout, cin)
        self.b = net.params[name + ".b"]
        self.cin, self.cout = cin, cout
        shadow = Param(name + ".w.gemm", (1, 1, 1, cin, 8 * cout), 0)
        shadow.w = torch.zeros(shadow.shape, dtype=torch.float32, device=DEV)
        shadow.grad = torch.zeros(shadow.shape, dtype=torch.float32, device=DEV)
        self.shadow = shadow
        self.gemm = Conv3D(net, name, 1, 1, cin, 8 * cout, use_bias=False, w_param=shadow)
        net.pre_pack.append(lambda: call("vg_conv3d_transpose_k2s2_weights", self.w.w, shadow.w, cin, cout, 0))

    def __call__(self, tape, x):
        n, d, h, w, c = x.shape
        inner = Tape(enabled=tape.enabled)
        t = self.gemm(inner, x)
        t.src = None                                       # no Var <-> Node cycle: the inner node is only reachable from bwd
        y = torch.empty((n, 2 * d, 2 * h, 2 * w, self.cout), dtype=torch.bfloat16, device=DEV)
        call("vg_conv3d_transpose_k2s2_scatter", t.data, self.b.w, y, n, d, h, w, self.cout)
        out = Var(y)

        def bwd(in_needs, p_needs):
            dt = torch.empty_like(t.data)
            call("vg_conv3d_transpose_k2s2_gather", out.grad, dt, self.b.grad if p_needs else None, n, d, h, w, self.cout)
            t.grad = dt
            if p_needs:
                self.shadow.grad.zero_()
            inner.nodes[0].bwd(in_needs, p_needs)          # pointwise wgrad (into the shadow) and dgrad (accumulates into x)
            t.grad = None
            if p_needs:
                call("vg_conv3d_transpose_k2s2_weights", self.shadow.grad, self.w.grad, self.cin, self.cout, 1)

        tape.record([x], [out], [self.w, self.b], bwd, "conv_transpose")
        return out


def upsample_concat(tape, lo, skip):
    """UpSampling3D(2) + concatenate([x, xskip]) (resunet_model.py:176,181)."""
    n, d, h, w, c0 = lo.shape
    c1 = skip.shape[-1]
    y = torch.empty((n, 2 * d, 2 * h, 2 * w, c0 + c1), dtype=torch.bfloat16, device=DEV)
    call("vg_upsample_concat", lo.data, skip.data, y, n, d, h, w, c0, c1)
    out = Var(y)
    # InstanceNorm statistics of the concat without reading it: nearest-neighbour upsampling repeats every value 8 times, so the
    # per-(n, c) mean and variance of the upsampled half are those of `lo` (1/8 of the voxels), and the skip half was normalised on
    # the encoder side already (its statistics are cached on the Var)
    if STATS_REUSE and skip.stats is not None:
        if lo.stats is None:
            ws_bytes = _lib.lib().vg_instnorm_workspace_bytes(n, d, h, w, c0)
            ws = torch.empty(ws_bytes // 4 + 1, dtype=torch.float32, device=DEV)
            m0 = torch.empty(n * c0, dtype=torch.float32, device=DEV)
            r0 = torch.empty(n * c0, dtype=torch.float32, device=DEV)
            call("vg_instnorm_stats", lo.data, dtype_code(lo.data), n, d, h, w, c0, m0, r0, ws, ws_bytes,
                 work=float(lo.data.element_size() * lo.data.numel()))
            lo.stats = (m0, r0)
        out.stats = tuple(torch.cat([a.view(n, c0), b.view(n, c1)], dim=1).reshape(-1).contiguous() for a, b in zip(lo.stats, skip.stats))

    def bwd(in_needs, p_needs):
        dlo = torch.empty_like(lo.data)
        fuse = in_needs[1] and skip.grad is not None     # add the skip gradient into the existing one in the same pass
        dsk = skip.grad if fuse else torch.empty_like(skip.data)
        call("vg_upsample_concat_bwd", out.grad, dlo, dsk, 1 if fuse else 0, n, d, h, w, c0, c1)
        if in_needs[0]:
            accumulate(lo, dlo)
        if in_needs[1] and not fuse:
            accumulate(skip, dsk)

    tape.record([lo, skip], [out], [], bwd, "upcat")
    return out


def pad_noise(tape, x, noise=None, noise_std=0.0, seed=0, seed_dev=None):
    """ReflectionPadding3D + GaussianNoise on a single-channel fp32 volume (discriminator.py:50-52)."""
    n, d, h, w, c = x.shape
    assert c == 1 and x.data.dtype == torch.float32
    y = torch.empty((n, d + 2, h + 2, w + 2, 1), dtype=torch.float32, device=DEV)
    call("vg_pad_noise", x.data, y, n, d, h, w, noise, noise_std if noise is None else 0.0, seed, seed_dev)
    out = Var(y)

    def bwd(in_needs, p_needs):
        if in_needs[0]:
            dx = torch.empty_like(x.data)
            call("vg_pad_fold", out.grad, dx, n, d, h, w, 0)
            accumulate(x, dx)

    tape.record([x], [out], [], bwd, "pad_noise")
    return out


def gather_pad(tape, a, b, up=1, pad=0, mode=PAD_ZERO):
    """out = pad(concat(upsample_up(a), b)) — UpSampling3D / concatenate / ReflectionPadding3D / TF-'same' zeros in one pass
    (vnet_model.py:116,132,247-252).  `a` or `b` may be None."""
    src = a if a is not None else b
    n = src.shape[0]
    if a is not None:
        d, h, w = [s * up for s in a.shape[1:4]]
    else:
        d, h, w = b.shape[1:4]
    c0 = a.shape[-1] if a is not None else 0
    c1 = b.shape[-1] if b is not None else 0
    y = torch.empty((n, d + 2 * pad, h + 2 * pad, w + 2 * pad, c0 + c1), dtype=torch.bfloat16, device=DEV)
    call("vg_gather_pad", a.data if a is not None else None, b.data if b is not None else None, y, n, d, h, w, c0, c1, up, pad, mode)
    out = Var(y)
    ins = [v for v in (a, b) if v is not None]

    def bwd(in_needs, p_needs):
        need = dict(zip([id(v) for v in ins], in_needs))
        da = torch.empty_like(a.data) if (a is not None and need[id(a)]) else None
        db = torch.empty_like(b.data) if (b is not None and need[id(b)]) else None
        call("vg_gather_pad_bwd", out.grad, da, db, n, d, h, w, c0, c1, up, pad, mode)
        if da is not None:
            accumulate(a, da)
        if db is not None:
            accumulate(b, db)

    tape.record(ins, [out], [], bwd, "gather_pad")
    return out


def upsample_pad(tape, x, lo, hi):
    """UpSampling3D(2) + the zero padding (lo before / hi after) of the 'same' convolution that follows (building_blocks.py:240-280)."""
    n, d, h, w, c = x.shape
    pp = lo + hi
    y = torch.empty((n, 2 * d + pp, 2 * h + pp, 2 * w + pp, c), dtype=torch.bfloat16, device=DEV)
    call("vg_upsample_pad", x.data, y, n, d, h, w, c, lo, hi)
    out = Var(y)

    def bwd(in_needs, p_needs):
        if in_needs[0]:
            dx = torch.empty_like(x.data)
            call("vg_upsample_pad_bwd", out.grad, dx, n, d, h, w, c, lo, hi)
            accumulate(x, dx)

    tape.record([x], [out], [], bwd, "upsample_pad")
    return out


def maxpool_pad(tape, x, pad=0, mode=PAD_ZERO):
    """pad(MaxPooling3D(2)(x)) (vnet_model.py:223 followed by the ReflectionPadding3D at :116)."""
    n, d, h, w, c = x.shape
    y = torch.empty((n, d // 2 + 2 * pad, h // 2 + 2 * pad, w // 2 + 2 * pad, c), dtype=torch.bfloat16, device=DEV)
    call("vg_maxpool2_pad", x.data, y, n, d, h, w, c, pad, mode)
    out = Var(y)

    def bwd(in_needs, p_needs):
        if in_needs[0]:
            dx = torch.empty_like(x.data)
            call("vg_maxpool2_pad_bwd", x.data, out.grad, dx, n, d, h, w, c, pad, mode)
            accumulate(x, dx)

    tape.record([x], [out], [], bwd, "maxpool_pad")
    return out


def dropout_mask(n, rate, seed, seed_dev=None):
    """SpatialDropout3D keep-mask / (1 - rate) for n (sample, channel) pairs, drawn on the device (Philox)."""
    out = torch.empty(n, dtype=torch.float32, device=DEV)
    call("vg_dropout_mask", out, n, float(rate), int(seed), seed_dev)
    return out
