"""VanGan — the reference's trainer class (vangan.py:20-508) re-backed by the B200 kernels.

Kept from the reference: the constructor signature, the attribute names the loss functions read,
`compute_losses`, `train_step`, `test_step`, `distributed_train_step`, `distributed_test_step`,
`reduce_dict`, the ten result-dict keys (vangan.py:338-351) and the order of the four optimizer
updates (vangan.py:426-438).  Built: the LSGAN branch with any of the three generator families ('resnet' -- the constructor's
default --, 'resUnet' -- what main.py:196-200 runs --, 'vnet'); the Wasserstein / gradient-penalty branch raises.
"""
import os

import numpy as np
import torch

from . import engine as E
from .discriminator import get_discriminator
from .distribute import Strategy
from .loss_functions import (LossContext, cycle_loss, cycle_reconstruction, cycle_seg_loss, discriminator_loss_fn,
                             generator_loss_fn)
from .generator import ResNetGenerator, get_resnet_generator
from .resunet_model import ResUNet
from .vnet_model import VNetModel, custom_vnet

RESULT_KEYS = ("total_IS_loss", "total_SI_loss", "D_I_loss", "D_S_loss", "gen_IS_loss", "gen_SI_loss",
               "cycle_gen_SIS_loss", "cycle_gen_ISI_loss", "seg_loss", "reconstruction_loss_I")


class VanGan:
    def __init__(self, args, strategy=None, lambda_cycle=10.0, lambda_identity=5, lambda_reconstruction=5,
                 lambda_topology=5, gen_i2s='resnet', gen_s2i='resnet', semi_supervised=False, wasserstein=False,
                 ncritic=5, gp_weight=10.0, seed=1234):
        self.strategy = strategy if strategy is not None else Strategy()
        self.n_devices = args.N_DEVICES
        self.img_size = args.INPUT_IMG_SIZE
        self.lambda_cycle = lambda_cycle
        self.lambda_identity = lambda_identity
        self.lambda_reconstruction = lambda_reconstruction
        self.lambda_topology = lambda_topology
        self.channels = args.CHANNELS
        self.gen_i2s_typ, self.gen_s2i_typ = gen_i2s, gen_s2i
        self.semi_supervised = semi_supervised
        self.global_batch_size = args.GLOBAL_BATCH_SIZE
        self.dims = args.DIMENSIONS
        if self.dims != 3:
            raise NotImplementedError("only DIMENSIONS=3 is built (main.py:80)")
        sp = args.SUBVOL_PATCH_SIZE
        self.subvol_patch_size = (sp[0], sp[1], sp[2], self.channels)
        self.seg_subvol_patch_size = (sp[0], sp[1], sp[2], 1)
        self.train_steps = getattr(args, "train_steps", None)
        self.batch_size = getattr(args, "BATCH_SIZE", None)
        self.wasserstein = wasserstein
        if wasserstein:
            raise NotImplementedError("Wasserstein/GP branch (vangan.py:355-378,400-423) is out of scope")
        self.cycle_loss_fn = cycle_loss
        self.generator_loss_fn = generator_loss_fn
        self.discriminator_loss_fn = discriminator_loss_fn
        self.seg_loss_fn = cycle_seg_loss
        self.reconstruction_loss = cycle_reconstruction
        self.layer_noise = 0.1
        self.cldice_iters = 15
        self.step = 0
        self.current_epoch = 0
        self.checkpoint_loaded = False     # read (and cleared) by GanMonitor.set_learning_rate; the reference never sets it either
        self.checkpoint_dir = os.path.join(getattr(args, "output_dir", "."), 'checkpoints')
        self.checkpoint_prefix = os.path.join(self.checkpoint_dir, 'checkpoint')
        self.loss_ctx = None
        self.tape = None
        self.last = None
        self.keep_last = False   # tests set this to inspect fake/cycled volumes after a step

        with self.strategy.scope():
            if gen_i2s not in ('resUnet', 'vnet', 'resnet') or gen_s2i not in ('resUnet', 'vnet', 'resnet'):
                raise ValueError('Generator type not recognised')      # vangan.py:124,164
            if gen_i2s == 'resnet':
                self.gen_IS = get_resnet_generator(input_img_size=self.subvol_patch_size, batch_size=self.global_batch_size,
                                                   name='generator_IS', num_downsampling_blocks=3, num_upsample_blocks=3, seed=seed)
            elif gen_i2s == 'vnet':
                self.gen_IS = custom_vnet(input_shape=self.subvol_patch_size, activation='relu', use_batch_norm=False,
                                          upsample_mode='upsample', dropout=0.5, dropout_change_per_layer=0.0,
                                          dropout_type='spatial', use_dropout_on_upsampling=False, use_attention_gate=False,
                                          filters=32, num_layers=4, output_activation='tanh', name='generator_IS', seed=seed)
            else:
                self.gen_IS = ResUNet(input_shape=self.subvol_patch_size, upsample_mode='simple', dropout=0.1,
                                      dropout_change_per_layer=0.1, dropout_type='none', use_attention_gate=False,
                                      filters=16, num_layers=4, name='generator_IS', seed=seed)
            if gen_s2i == 'resnet':
                self.gen_SI = get_resnet_generator(input_img_size=self.subvol_patch_size, batch_size=self.global_batch_size,
                                                   name='generator_SI', num_downsampling_blocks=3, num_upsample_blocks=3, seed=seed + 1)
            elif gen_s2i == 'vnet':
                self.gen_SI = custom_vnet(input_shape=self.subvol_patch_size, activation='relu', use_batch_norm=True,
                                          upsample_mode='deconv', dropout=0.5, dropout_change_per_layer=0.0,
                                          dropout_type='spatial', use_dropout_on_upsampling=False, use_attention_gate=False,
                                          filters=16, num_layers=4, output_activation='tanh', addnoise=False,
                                          name='generator_SI', seed=seed + 1)
            else:
                self.gen_SI = ResUNet(input_shape=self.seg_subvol_patch_size, upsample_mode='simple', dropout=0.1,
                                      dropout_change_per_layer=0.1, dropout_type='none', use_attention_gate=False,
                                      filters=16, num_layers=4, use_input_noise=False, name='generator_SI', seed=seed + 1)
            common = dict(filters=64, use_dropout=True, dropout_rate=0.2, wasserstein=False, use_SN=False,
                          use_input_noise=True, use_layer_noise=True, noise_std=self.layer_noise)
            self.disc_I = get_discriminator(input_img_size=self.subvol_patch_size, batch_size=self.global_batch_size,
                                            name='discriminator_I', seed=seed + 2, **common)
            self.disc_S = get_discriminator(input_img_size=self.seg_subvol_patch_size, batch_size=self.global_batch_size,
                                            name='discriminator_S', seed=seed + 3, **common)
            # tf.keras.optimizers.Adam(2e-4, beta_1=0.5, beta_2=0.9, clipnorm=100) x4 (vangan.py:220-235)
            mk = lambda net: E.Adam(net, learning_rate=2e-4, beta_1=0.5, beta_2=0.9, epsilon=1e-7, clipnorm=100.0)
            self.gen_I_optimizer, self.gen_S_optimizer = mk(self.gen_IS), mk(self.gen_SI)
            self.disc_I_optimizer, self.disc_S_optimizer = mk(self.disc_I), mk(self.disc_S)
        self.networks = {"gen_IS": self.gen_IS, "gen_SI": self.gen_SI, "disc_I": self.disc_I, "disc_S": self.disc_S}
        self.optimizers = {"gen_I_optimizer": self.gen_I_optimizer, "gen_S_optimizer": self.gen_S_optimizer,
                           "disc_I_optimizer": self.disc_I_optimizer, "disc_S_optimizer": self.disc_S_optimizer}
        # per-step state that lives in DEVICE memory so that a captured CUDA graph of the whole step can be replayed:
        # the RNG seed offset of the discriminators' GaussianNoise / SpatialDropout3D and the four Adam step sizes lr_t
        self._seed_dev = torch.zeros(1, dtype=torch.int64, device=E.DEV)
        self._lr_dev = torch.zeros(4, dtype=torch.float32, device=E.DEV)
        self._h_seed = torch.zeros(1, dtype=torch.int64).pin_memory()
        self._h_lr = torch.zeros(4, dtype=torch.float32).pin_memory()
        self.seed = seed
        # independent branches of the step (the two generator chains of the forward pass, the four backward sweeps) are enqueued on
        # side streams forked from / joined to the caller's stream by events, so small kernels of one branch fill the SMs another
        # leaves idle (per-GPU batch 1 on 8 GPUs); VG_STREAMS=0 keeps everything on one stream
        self.use_streams = os.environ.get("VG_STREAMS", "1") != "0"
        self._side = [torch.cuda.Stream() for _ in range(8)] if self.use_streams else None
        # one more stream per sweep for the weight-gradient kernels (VG_WG_STREAM=0 disables)
        self._wg_side = [torch.cuda.Stream() for _ in range(8)] if (self.use_streams and os.environ.get("VG_WG_STREAM", "1") != "0") else None
        # VG_SPLIT_SWEEPS=1: every loss swept as its independent parts (8 sweeps side by side).  Measured (call 38): b=8 175.0 -> 176.6
        # ms, b=1 32.3 -> 31.5 ms -- the SMs are already busy with four sweeps, and the extra concurrent atomics widen the run-to-run
        # spread of the gradients -- so the default stays one sweep per network.
        self.split_sweeps = self.use_streams and os.environ.get("VG_SPLIT_SWEEPS", "0") == "1"
        # CUDA graph of one full train step (captured on the third eligible call; VG_GRAPH=0 disables)
        self.use_graph = os.environ.get("VG_GRAPH", "1") != "0" and not isinstance(self.gen_IS, VNetModel) and not isinstance(self.gen_SI, VNetModel)
        self._graph = None
        self._eager_steps = 0
        self.launches_per_replay = 0

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _as_var(x):
        if isinstance(x, E.Var):
            return x
        t = torch.as_tensor(x)
        if t.device.type != "cuda":
            t = t.to(E.DEV, non_blocking=True)
        return E.Var(t.to(torch.float32).contiguous())

    def _gen(self, net, tape, x, training, app):
        """One generator application; the V-Net variant draws its SpatialDropout3D masks per (step, application)."""
        if isinstance(net, VNetModel):
            return net.forward(tape, x, training=training, seed=self.step * 4 + app)
        if isinstance(net, ResNetGenerator):      # SpatialDropout3D masks: per-application key + the per-step device offset (graph-safe)
            return net.forward(tape, x, training=training, seed=4 + app, seed_dev=self._seed_dev)
        return net.forward(tape, x)

    def _disc(self, net, tape, x, training, rand, key, app):
        noise, masks = (None, None) if rand is None else rand[key]
        # the seed is a per-application constant; the per-step part is the device-resident offset self._seed_dev
        return net.forward(tape, x, training=training, noise=noise, masks=masks, seed=app, seed_dev=self._seed_dev)

    # ------------------------------------------------------------------ reference API
    def compute_losses(self, real_I, real_S, result, training=True, rand=None, tape=None):
        """vangan.py:270-353.  Returns (result, total_loss_I, total_loss_S, disc_I_loss, disc_S_loss, fake_I, fake_S);
        the losses are `Scalar`s (float() for the value).  `rand`: optional explicit discriminator noise / dropout
        tensors {'S_real','S_fake','I_real','I_fake': (noise list, mask list)} for parity tests."""
        tape = tape if tape is not None else E.Tape(enabled=training)
        self.tape = tape
        self.loss_ctx = LossContext()
        real_I, real_S = self._as_var(real_I), self._as_var(real_S)
        # Four chains on side streams: A = gen_IS applications + the S-cycle losses, B = gen_SI applications + the I-cycle losses,
        # C = disc_S (real, then fake), D = disc_I.  They meet only where one consumes another's generated volume.
        main = torch.cuda.current_stream()
        par = self._side is not None and rand is None
        sA, sB, sC, sD = self._side[:4] if par else (main, main, main, main)
        if par:
            self._fork(main, self._side[:4])
        with torch.cuda.stream(sA):
            fake_S = self._gen(self.gen_IS, tape, real_I, training, 0)
        with torch.cuda.stream(sB):
            fake_I = self._gen(self.gen_SI, tape, real_S, training, 1)
        with torch.cuda.stream(sC):
            disc_real_S = self._disc(self.disc_S, tape, real_S, training, rand, "S_real", 0)
        with torch.cuda.stream(sD):
            disc_real_I = self._disc(self.disc_I, tape, real_I, training, rand, "I_real", 2)
        if par:
            eA, eB = torch.cuda.Event(), torch.cuda.Event()
            eA.record(sA); eB.record(sB)
            sA.wait_event(eB); sB.wait_event(eA)      # cycled_S needs fake_I, cycled_I needs fake_S
            sC.wait_event(eA); sD.wait_event(eB)      # disc_S(fake_S), disc_I(fake_I)
        with torch.cuda.stream(sA):
            cycled_S = self._gen(self.gen_IS, tape, fake_I, training, 2)
            cycle_loss_I = self.cycle_loss_fn(self, real_S, cycled_S, typ="bce")
            seg_loss = self.seg_loss_fn(self, real_S, cycled_S, iters=self.cldice_iters)
        with torch.cuda.stream(sB):
            cycled_I = self._gen(self.gen_SI, tape, fake_S, training, 3)
            cycle_loss_S = self.cycle_loss_fn(self, real_I, cycled_I, typ='mse')
            reconstruction_loss = self.reconstruction_loss(self, real_I, cycled_I)
        with torch.cuda.stream(sC):
            disc_fake_S = self._disc(self.disc_S, tape, fake_S, training, rand, "S_fake", 1)
            gen_IS_loss = self.generator_loss_fn(self, disc_fake_S, from_logits=True)
            disc_S_loss = self.discriminator_loss_fn(self, disc_real_S, disc_fake_S, from_logits=True)
        with torch.cuda.stream(sD):
            disc_fake_I = self._disc(self.disc_I, tape, fake_I, training, rand, "I_fake", 3)
            gen_SI_loss = self.generator_loss_fn(self, disc_fake_I, from_logits=True)
            disc_I_loss = self.discriminator_loss_fn(self, disc_real_I, disc_fake_I, from_logits=True)
        if par:
            self._join(main, self._side[:4])

        total_loss_I = gen_IS_loss + cycle_loss_I + seg_loss
        total_loss_S = gen_SI_loss + cycle_loss_S + reconstruction_loss
        result.update({
            'total_IS_loss': total_loss_I, 'total_SI_loss': total_loss_S, 'D_I_loss': disc_I_loss,
            'D_S_loss': disc_S_loss, 'gen_IS_loss': gen_IS_loss, 'gen_SI_loss': gen_SI_loss,
            'cycle_gen_SIS_loss': cycle_loss_I, 'cycle_gen_ISI_loss': cycle_loss_S, 'seg_loss': seg_loss,
            'reconstruction_loss_I': reconstruction_loss})
        self.last = dict(fake_S=fake_S, fake_I=fake_I, cycled_S=cycled_S, cycled_I=cycled_I, disc_real_S=disc_real_S,
                         disc_fake_S=disc_fake_S, disc_real_I=disc_real_I, disc_fake_I=disc_fake_I)
        return result, total_loss_I, total_loss_S, disc_I_loss, disc_S_loss, fake_I, fake_S

    @staticmethod
    def _fork(main, sides):
        ev = torch.cuda.Event()
        ev.record(main)
        for st in sides:
            st.wait_event(ev)

    @staticmethod
    def _join(main, sides):
        for st in sides:
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)

    def _sweeps(self, pairs, overlap_allreduce):
        """Backward sweeps of `pairs` = [(net, [loss part, ...]), ...].  Every PART is swept on its own side stream: the parts of one
        loss seed different tensors (a generator's adversarial term reaches it through the discriminator and its first application, its
        cycle terms through the second application; a discriminator's real and fake terms through its two applications), backward is
        linear in the seeds, parts share only read-only activations and weights, and every parameter-gradient kernel accumulates with
        atomics -- so they run side by side.  A network's all-reduce is enqueued once all of its parts are done."""
        main = torch.cuda.current_stream()
        nparts = sum(len(parts) for _net, parts in pairs)
        sides = self._side[:nparts] if (self._side is not None and nparts > 1) else None
        handles = []
        for net, _parts in pairs:
            net.zero_grad()
        if sides is not None:
            self._fork(main, sides)
        k = 0
        for net, parts in pairs:
            first = k
            for part in parts:
                with torch.cuda.stream(sides[k] if sides is not None else main):
                    self.tape.wg_stream = self._wg_side[k] if (self._wg_side is not None and sides is not None) else None
                    self.tape.backward(part.seeds(), net.trainable_variables)
                    self.tape.wg_stream = None
                k += 1
            with torch.cuda.stream(sides[first] if sides is not None else main):
                if sides is not None and k - first > 1:
                    self._join(torch.cuda.current_stream(), sides[first + 1:k])
                # MirroredStrategy's gradient all-reduce: enqueued on the communication stream as soon as this network's sweeps end
                handles.append(self.strategy.all_reduce_async(net.g) if overlap_allreduce else None)
        if sides is not None:
            self._join(main, sides)
        return handles

    def _plan(self, total_I, total_S, dI, dS):
        """(network, parts of the loss its `minimize` differentiates) in the reference's order (vangan.py:426-438).  total_loss_I =
        gen_IS_loss + cycle_loss_I + seg_loss: the first term seeds disc_S(fake_S), the other two seed cycled_S; the discriminator
        losses are MSE(1, D(real)) + MSE(0, D(fake)), one term per application."""
        if not self.split_sweeps:
            return ((self.gen_IS, [total_I]), (self.gen_SI, [total_S]), (self.disc_I, [dI]), (self.disc_S, [dS]))
        assert len(total_I.grad_fns) == 3 and len(total_S.grad_fns) == 3 and len(dI.grad_fns) == 2 and len(dS.grad_fns) == 2
        return ((self.gen_IS, [total_I.part(0), total_I.part(1, 2)]), (self.gen_SI, [total_S.part(0), total_S.part(1, 2)]),
                (self.disc_I, [dI.part(0), dI.part(1)]), (self.disc_S, [dS.part(0), dS.part(1)]))

    @staticmethod
    def seed_offset(seed, step, world=1, rank=0):
        """Offset added to every in-kernel Philox key of one step.  A (step, replica) pair owns a block of 128 keys (the four
        discriminator applications use key = application * 16 + layer, the 'resnet' generator applications (4 + application) * 16 +
        dropout layer), so draws differ per step AND per replica --
        MirroredStrategy draws GaussianNoise / SpatialDropout3D independently on every replica -- while `seed` itself, which
        also initialises the weights, stays shared across ranks."""
        return ((seed * 1000003 + step) * world + rank) * 128

    def _upload_step_state(self):
        """Seed offset and the four Adam step sizes of THIS step -> device (async copies on the current stream)."""
        self._h_seed[0] = self.seed_offset(self.seed, self.step, self.strategy.num_replicas_in_sync, self.strategy.rank)
        for i, opt in enumerate(self.optimizers.values()):
            self._h_lr[i] = opt.step_size()
        self._seed_dev.copy_(self._h_seed, non_blocking=True)
        self._lr_dev.copy_(self._h_lr, non_blocking=True)

    def _body_losses_and_sweeps(self, real_I, real_S, rand, overlap_allreduce):
        """compute_losses + the four backward sweeps (vangan.py:394-438), enqueue only.  Returns (result, plan, handles)."""
        result = {}
        result, total_I, total_S, dI, dS, _fI, _fS = self.compute_losses(real_I, real_S, result, training=True, rand=rand)
        plan = self._plan(total_I, total_S, dI, dS)
        handles = self._sweeps(plan, overlap_allreduce)
        return result, plan, handles

    def _body_adam(self, only=None):
        """clip + Adam + operand repack of every network, or of those in `only` (the generators' update runs beside the
        discriminators' all-reduce in the multi-GPU layout)"""
        for i, opt in enumerate(self.optimizers.values()):
            if only is not None and opt.net not in only:
                continue
            opt.net.adam_step(lr_t_dev=self._lr_dev[i:i + 1], beta_1=opt.beta_1, beta_2=opt.beta_2, eps=opt.epsilon,
                              clipnorm=opt.clipnorm)

    def _step_body(self, real_I, real_S, rand, apply):
        """One eager step: no host synchronisation inside."""
        result, plan, handles = self._body_losses_and_sweeps(real_I, real_S, rand, True)
        for h in handles:
            if h is not None:
                h.wait()
        if apply:
            self._body_adam()
        return result

    def _finish_step(self, result, ctx, apply=True):
        if apply:
            for net in self.networks.values():
                net.step_count += 1
        self.step += 1
        vals = ctx.values()            # the one device->host read of the step
        return {k: float(v.value_fn(vals)) for k, v in result.items()}

    def train_step(self, real_I, real_S, rand=None, apply=True):
        """vangan.py:380-440: persistent tape around compute_losses, then one minimize per network
        (gen_IS on total_loss_I, gen_SI on total_loss_S, disc_I, disc_S).  Each network's gradient
        all-reduce (MirroredStrategy's, hidden inside `minimize`) is launched as soon as its sweep ends.

        The whole step is enqueue-only (no host sync before the final read of the ten loss sums), so after two eager
        steps it is captured ONCE into a CUDA graph and replayed: ~2 000 kernel launches become one graph launch, which is
        what keeps a GPU busy when the per-GPU batch is 1 (8-GPU strong scaling).  `rand` (explicit noise tensors, parity
        tests) and `apply=False` always run eagerly."""
        graphable = self.use_graph and rand is None and apply
        if self._graph is not None and self._graph["noise"] != (self.disc_I.noise_std, self.disc_S.noise_std):
            self._graph = None      # GaussianNoise stddev is a launch constant of the captured kernels: re-capture (once per epoch,
            #                         GanMonitor.updateDiscriminatorNoise)
        if graphable and self._graph is not None and self._graph["key"] == self._shape_key(real_I, real_S):
            return self._replay(real_I, real_S)
        if graphable and self._eager_steps >= 2 and self._graph is None:
            try:
                self._capture(real_I, real_S)
                return self._replay(real_I, real_S)
            except Exception as e:      # capture is an optimisation: fall back to eager launches, loudly
                import warnings
                warnings.warn("CUDA graph capture of train_step failed (%s: %s); running eagerly" % (type(e).__name__, e))
                self.use_graph = False
                self._graph = None
                torch.cuda.synchronize()
        self._upload_step_state()
        result = self._step_body(real_I, real_S, rand, apply)
        out = self._finish_step(result, self.loss_ctx, apply)
        self._eager_steps += 1
        self._release()
        return out

    @staticmethod
    def _shape_key(real_I, real_S):
        a = real_I.data if isinstance(real_I, E.Var) else real_I
        b = real_S.data if isinstance(real_S, E.Var) else real_S
        return (tuple(a.shape), tuple(b.shape))

    def _capture(self, real_I, real_S):
        """The step as CUDA graphs.
        World 1: ONE graph -- losses, the four backward sweeps, clip+Adam, operand repack.
        World > 1: graph 1 = forward pass, losses and the two generator sweeps; graph 2 = the two discriminator sweeps; graphs 3 / 4 =
        clip+Adam of the generators / of the discriminators.  Between the replays the bucketed gradient all-reduces of the networks
        just swept are enqueued on the communication stream (vg_comm, ordered by events): the generators' messages run while the
        discriminator sweeps execute, the discriminators' while the generators are updated, and each update graph is ordered after
        its own messages.  Replays and collectives are all asynchronous: the host enqueues the whole step without waiting.
        VG_GRAPH_COMM=1 captures the collectives INTO a single graph instead (measured: fine at 32^3, hangs at 4x128^3 per GPU on
        2 GPUs -- kept opt-in for investigation)."""
        from . import _lib
        key = self._shape_key(real_I, real_S)
        gI = torch.empty(key[0], dtype=torch.float32, device=E.DEV)
        gS = torch.empty(key[1], dtype=torch.float32, device=E.DEV)
        world = self.strategy.num_replicas_in_sync
        in_graph_comm = world > 1 and os.environ.get("VG_GRAPH_COMM", "0") == "1" and self.strategy._ensure_comm()
        torch.cuda.synchronize()
        l0 = _lib.lib().vg_launch_count()
        graphs = []

        def capture(fn):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=graphs[0].pool() if graphs else None, capture_error_mode="thread_local"):
                out = fn()
            graphs.append(g)
            return out

        gsplit = os.environ.get("VG_GRAPH_SPLIT", "0")      # "1" / "2": the multi-GPU capture layouts at world 1 (tests)
        split = (world > 1 and not in_graph_comm) or gsplit in ("1", "2")
        # multi-GPU layout: "per-sweep" (four graphs: generator sweeps / discriminator sweeps / the two updates, with the generators'
        # all-reduces beside the discriminator sweeps and the discriminators' beside the generators' update) or "two" (forward +
        # all four sweeps side by side / clip+Adam, every all-reduce between the two: full four-way concurrency, exchange exposed)
        layout = "two" if (gsplit == "2" or (gsplit != "1" and os.environ.get("VG_GRAPH_LAYOUT", "per-sweep") == "two")) else "per-sweep"
        if not split:
            def whole():
                result, _plan, handles = self._body_losses_and_sweeps(E.Var(gI), E.Var(gS), None, in_graph_comm)
                for h in handles:
                    if h is not None:
                        h.wait()
                self._body_adam()
                return result
            result = capture(whole)
            mode = "single"
        elif layout == "two":
            result = capture(lambda: self._body_losses_and_sweeps(E.Var(gI), E.Var(gS), None, False)[0])
            capture(self._body_adam)
            mode = "two"
        else:
            state = {}

            def first():
                res = {}
                res, total_I, total_S, dI, dS, _fI, _fS = self.compute_losses(E.Var(gI), E.Var(gS), res, training=True, rand=None)
                state["plan"] = self._plan(total_I, total_S, dI, dS)
                self._sweeps(state["plan"][:2], False)          # both generator sweeps, side by side
                return res
            result = capture(first)
            capture(lambda: self._sweeps(state["plan"][2:], False))   # both discriminator sweeps
            # two update graphs: the generators' (their all-reduces finished beside the discriminator sweeps) runs while the
            # discriminators' messages are in flight, the discriminators' after them
            capture(lambda: self._body_adam(only=(self.gen_IS, self.gen_SI)))
            capture(lambda: self._body_adam(only=(self.disc_I, self.disc_S)))
            mode = "per-sweep"
        self.launches_per_replay = int(_lib.lib().vg_launch_count() - l0)
        ctx = self.loss_ctx
        # the graphs' private pool keeps every buffer the capture touched; the Python-side tape is not needed again
        self.tape.clear()
        self.tape, self.last = None, None
        self._graph = dict(key=key, graphs=graphs, mode=mode, I=gI, S=gS, result=result, ctx=ctx,
                           noise=(self.disc_I.noise_std, self.disc_S.noise_std))

    def _sweep(self, net, loss):
        net.zero_grad()
        self.tape.backward(loss.seeds(), net.trainable_variables)

    def backward_of(self, net, loss):
        """one sweep of `loss` w.r.t. the variables of `net` on the current stream (tests)"""
        self._sweep(net, loss)

    def _replay(self, real_I, real_S):
        g = self._graph
        a = real_I.data if isinstance(real_I, E.Var) else torch.as_tensor(real_I)
        b = real_S.data if isinstance(real_S, E.Var) else torch.as_tensor(real_S)
        g["I"].copy_(a, non_blocking=True)        # H2D (pinned host batch) or D2D
        g["S"].copy_(b, non_blocking=True)
        self._upload_step_state()
        if g["mode"] == "single":
            g["graphs"][0].replay()
        elif g["mode"] == "two":
            g["graphs"][0].replay()
            handles = [self.strategy.all_reduce_async(net.g) for net in (self.gen_IS, self.gen_SI, self.disc_I, self.disc_S)]
            for h in handles:
                if h is not None:
                    h.wait()
            g["graphs"][1].replay()
        else:
            gens, discs = (self.gen_IS, self.gen_SI), (self.disc_I, self.disc_S)
            g["graphs"][0].replay()                                                   # forward + generator sweeps
            gen_h = [self.strategy.all_reduce_async(net.g) for net in gens]           # ... their messages run beside graph 1
            g["graphs"][1].replay()                                                   # discriminator sweeps
            disc_h = [self.strategy.all_reduce_async(net.g) for net in discs]
            for h in gen_h:
                if h is not None:
                    h.wait()
            g["graphs"][2].replay()                                                   # clip+Adam of the generators, beside the discriminators' messages
            for h in disc_h:
                if h is not None:
                    h.wait()
            g["graphs"][3].replay()                                                   # clip+Adam of the discriminators
        g["ctx"].host = None
        return self._finish_step(g["result"], g["ctx"], True)

    def _release(self):
        """Drop the step's tape, loss scratch and (unless keep_last) the generated volumes."""
        if self.tape is not None:
            self.tape.clear()
        self.tape = None
        self.loss_ctx = None
        if not self.keep_last:
            self.last = None

    def test_step(self, real_I, real_S):
        """vangan.py:442-457."""
        result = {}
        result, *_ = self.compute_losses(real_I, real_S, result, training=False)
        vals = self.loss_ctx.values()
        out = {k: float(v.value_fn(vals)) for k, v in result.items()}
        self._release()
        return out

    def reduce_dict(self, d):
        """vangan.py:459-473: SUM over replicas of every entry (one 10-float all-reduce)."""
        keys = list(d.keys())
        if self.strategy.num_replicas_in_sync > 1:
            t = torch.tensor([d[k] for k in keys], dtype=torch.float64, device=E.DEV)
            self.strategy.reduce("SUM", t, axis=None)
            vals = t.cpu().tolist()
            for k, v in zip(keys, vals):
                d[k] = v

    def distributed_train_step(self, x, y, rand=None):
        """vangan.py:475-490.  x, y: this replica's shard of the global batch."""
        results = self.strategy.run(self.train_step, args=(x, y), kwargs=dict(rand=rand))
        self.reduce_dict(results)
        return results

    def distributed_test_step(self, x, y):
        results = self.strategy.run(self.test_step, args=(x, y))
        self.reduce_dict(results)
        return results

    # ------------------------------------------------------------------ checkpoints
    # tf.train.Checkpoint over the four models and the four optimizers (vangan.py:238-268).  TensorFlow's on-disk format needs
    # TensorFlow; the same object graph is written as ONE .npz per epoch whose keys follow the Checkpoint's attribute names:
    #   <model>/<variable>            fp32, Keras layout (Conv3D kernel (kd,kh,kw,Cin,Cout), bias, gamma, beta)
    #   <optimizer>/m/<variable>, <optimizer>/v/<variable>     Adam slots
    #   <optimizer>/iter              Keras `optimizer.iterations`
    #   save_counter/step             the trainer's step counter (keys the in-kernel noise / dropout streams)
    _OPT_OF = {"gen_IS": "gen_I_optimizer", "gen_SI": "gen_S_optimizer", "disc_I": "disc_I_optimizer", "disc_S": "disc_S_optimizer"}

    def _checkpoint_path(self, epoch):
        return self.checkpoint_prefix + "_e{epoch}".format(epoch=epoch) + ".npz"

    def save_checkpoint(self, epoch):
        """vangan.py:247-250: writes <output_dir>/checkpoints/checkpoint_e{epoch+1} (overwrites)."""
        os.makedirs(os.path.dirname(self.checkpoint_prefix), exist_ok=True)
        arrays = {"save_counter/step": np.int64(self.step)}
        for nn, net in self.networks.items():
            on = self._OPT_OF[nn]
            m, v = net.m.cpu().numpy(), net.v.cpu().numpy()
            for k, val in net.export().items():
                p = net.params[k]
                arrays[nn + "/" + k] = val
                arrays[on + "/m/" + k] = m[p.offset:p.offset + p.size].reshape(p.shape)
                arrays[on + "/v/" + k] = v[p.offset:p.offset + p.size].reshape(p.shape)
            arrays[on + "/iter"] = np.int64(net.step_count)
            for k, buf in net.buffers.items():          # non-trainable variables (BatchNormalization moving statistics)
                arrays[nn + "/" + k] = buf.cpu().numpy()
        path = self._checkpoint_path(epoch + 1)
        np.savez(path, **arrays)
        print(f'\nSaved checkpoint to {self.checkpoint_prefix}\n')
        return path

    def load_checkpoint(self, epoch=None, expect_partial: bool = False, newpath=None):
        """vangan.py:252-268: restores models AND optimizers from checkpoint_e{epoch} (under `newpath` when given).  Prints the
        reference's error line when the file is missing.  expect_partial: tolerate missing optimizer slots."""
        if newpath is not None:
            self.checkpoint_prefix = os.path.join(newpath, 'checkpoint')
        path = self._checkpoint_path(epoch)
        print(f"Trying to load checkpoint from path: {path}")
        if not os.path.exists(path):
            print('Error: Checkpoint not found!')
            return False
        z = np.load(path)
        for nn, net in self.networks.items():
            on = self._OPT_OF[nn]
            net.load({k: z[nn + "/" + k] for k in net.params})
            for k, buf in net.buffers.items():
                buf.copy_(torch.from_numpy(np.ascontiguousarray(z[nn + "/" + k])).to(E.DEV))
            have_slots = all((on + "/m/" + k) in z.files for k in net.params)
            if not have_slots and not expect_partial:
                raise KeyError("checkpoint %s has no optimizer slots for %s (pass expect_partial=True)" % (path, on))
            if have_slots:
                for k, p in net.params.items():
                    net.m[p.offset:p.offset + p.size].copy_(torch.from_numpy(np.ascontiguousarray(z[on + "/m/" + k]).reshape(-1)).to(E.DEV))
                    net.v[p.offset:p.offset + p.size].copy_(torch.from_numpy(np.ascontiguousarray(z[on + "/v/" + k]).reshape(-1)).to(E.DEV))
                net.step_count = int(z[on + "/iter"])
        if "save_counter/step" in z.files:
            self.step = int(z["save_counter/step"])
        print(f'Loaded checkpoint from {path}\n')
        return True


def train(ds, gan, summary, epoch, steps=None, desc=None, training=True):
    """vangan.py:510-550: loops the dataset, appends the per-step result dicts."""
    from .utils import append_dict
    results, cntr = {}, 0
    for x, y in ds:
        if cntr == steps:
            break
        cntr += 1
        result = gan.distributed_train_step(x, y) if training else gan.distributed_test_step(x, y)
        append_dict(results, result)
    if summary is not None:
        for key, value in results.items():
            summary.scalar(key, float(np.mean(value)), epoch=epoch, training=training)
    return results
