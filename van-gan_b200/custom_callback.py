"""GanMonitor — the reference's sliding-window predictor (custom_callback.py:47-223, 466-509) with the
window loop, accumulation, overlap-count division and min-max rescale done on the device.

Kept from the reference: `stitch_subvolumes(gen, img, subvol_size, epoch, stride, name, output_path,
complete, padFactor, border_removal, process_img)` and `run_mapping(model, test_set, sub_img_size,
segmentation, stride, padFactor, filetext, filepath)`, the exact window enumeration (dim_out+1
iterations per axis with end clamping, so the last window is duplicated when (dim-k) is a multiple
of the stride), the 10 % border crop, UNIFORM overlap-count blending and the final
`255 * min_max_norm`.  Changed: windows are run through the generator in batches, volumes stay in
HBM, windows are sharded round-robin over the data-parallel ranks (one all-reduce of the partial
sums at the end).  The TIFF writers (skimage) are out of scope: results are returned, and saved as
.npy when an output path is given.
"""
import os

import numpy as np
import torch

from . import engine as E
from ._lib import call
from .distribute import Strategy


class _GraphedForward:
    """One generator forward for a fixed window-batch shape, captured into a CUDA graph and replayed per batch: the ~150 kernel
    launches of a forward cost more host time than device time at batch 4 (the serial loop of custom_callback.py:174-175 is
    launch-bound on a B200)."""

    def __init__(self, gen, shape):
        self.x = torch.zeros(shape, dtype=torch.float32, device=E.DEV)
        for _ in range(2):                      # warm-up outside the capture (allocator, lazy kernel attributes)
            gen(self.x, training=False)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.y = gen(self.x, training=False)

    def replay(self):
        self.graph.replay()
        return self.y


def _apply_hook(hook, win, batched):
    """The reference's `process_imaging_domain` hook (main.py:169-177 passes `process_imaging_otf`).  The stitcher calls it per
    window as hook(arr, axis=None, keepdims=False) (custom_callback.py:171-172), the plotter as hook(sample) on a batch of one
    (:262-263).  Hooks marked `_vg_device` (utils.process_imaging_otf) run on the device on the whole window batch -- per-sample
    reduction over a batch of windows is the same arithmetic as axis=None on each window; any other callable gets each window as
    a host numpy array, like the reference."""
    if getattr(hook, "_vg_device", False):
        return hook(win, axis=(1, 2, 3, 4), keepdims=True)
    out = []
    for i in range(win.shape[0]):
        a = win[i].cpu().numpy()
        r = hook(a[None]) if batched else hook(a, axis=None, keepdims=False)
        r = np.asarray(r, dtype=np.float32)
        out.append(r[0] if batched else r)
    return torch.as_tensor(np.stack(out), dtype=torch.float32, device=E.DEV).contiguous()


def window_starts(n, k, s):
    """custom_callback.py:127-162: dim_out+1 iterations, start clamped to n-k."""
    dim_out = int(np.floor((n - k) / s + 1))
    out, start = [], 0
    for _ in range(dim_out + 1):
        if start > n - k:
            start = n - k
        out.append(start)
        start += s
    return out


class GanMonitor:
    def __init__(self, args=None, dataset=None, imaging_val_data=None, segmentation_val_data=None, process_imaging_domain=None,
                 strategy=None, window_batch=4):
        """Same leading arguments as the reference (custom_callback.py:15-31); `strategy` / `window_batch` are this package's."""
        self.imgSize = getattr(args, "INPUT_IMG_SIZE", None)
        self.dims = getattr(args, "DIMENSIONS", 3) if args is not None else 3
        if self.dims != 3:
            raise NotImplementedError("only DIMENSIONS=3 is built (main.py:80)")
        self.imaging_val_full_vol_data = getattr(dataset, "imaging_val_full_vol_data", None)
        self.segmentation_val_full_vol_data = getattr(dataset, "segmentation_val_full_vol_data", None)
        self.imaging_val_data = imaging_val_data
        self.segmentation_val_data = segmentation_val_data
        self.process_imaging_domain = process_imaging_domain
        self.period = getattr(args, "PERIOD_2D_CALLBACK", 2)
        self.period3D = getattr(args, "PERIOD_3D_CALLBACK", 2)
        self.model_path = getattr(args, "output_dir", None)
        self.strategy = strategy if strategy is not None else Strategy()
        self.window_batch = window_batch
        self.use_graph = os.environ.get("VG_GRAPH", "1") != "0"
        self.last_stats = None
        self.last_panels = None
        self._fwd_cache = {}

    # ------------------------------------------------------------------ epoch callbacks (custom_callback.py:33-45,326-464)
    def save_model(self, model, epoch):
        """custom_callback.py:33-45: one export per network under <output_dir>/checkpoints/e{epoch+1}_{genAB,genBA,discA,discB}
        (Keras SavedModel in the reference; here the variables in Keras layout as .npz)."""
        d = os.path.join(self.model_path, "checkpoints")
        os.makedirs(d, exist_ok=True)
        for net, tag in ((model.gen_IS, "genAB"), (model.gen_SI, "genBA"), (model.disc_I, "discA"), (model.disc_S, "discB")):
            np.savez(os.path.join(d, "e{epoch}_{tag}.npz".format(epoch=epoch + 1, tag=tag)), **net.export())

    def set_learning_rate(self, model, epoch, args):
        """custom_callback.py:326-397: at epoch == INITIATE_LR_DECAY every optimizer's lr becomes a linear PolynomialDecay to 0
        over the remaining steps; after a checkpoint load past that epoch the schedule is re-derived from the epoch."""
        opts = (model.gen_I_optimizer, model.gen_S_optimizer, model.disc_I_optimizer, model.disc_S_optimizer)
        if epoch == args.INITIATE_LR_DECAY:
            for o in opts:
                o.lr = E.PolynomialDecay(initial_learning_rate=args.INITIAL_LR,
                                         decay_steps=(args.EPOCHS - args.INITIATE_LR_DECAY) * args.train_steps,
                                         end_learning_rate=0, power=1)
        if model.checkpoint_loaded and epoch > args.INITIATE_LR_DECAY:
            model.checkpoint_loaded = False
            learning_gradient = args.INITIAL_LR / (args.EPOCHS - args.INITIATE_LR_DECAY)
            intermediate_learning_rate = learning_gradient * (args.EPOCHS - epoch)
            print('Initial learning rate: %0.8f' % intermediate_learning_rate)
            for o in opts:
                o.lr = E.PolynomialDecay(initial_learning_rate=intermediate_learning_rate,
                                         decay_steps=(args.EPOCHS - args.INITIATE_LR_DECAY - epoch) * args.train_steps,
                                         end_learning_rate=0, power=1)

    def updateDiscriminatorNoise(self, model, init_noise, epoch, args):
        """custom_callback.py:399-424: stddev of every GaussianNoise layer of one discriminator decays linearly to 0 at NO_NOISE."""
        decay_rate = 1. if args.NO_NOISE == 0 else epoch / args.NO_NOISE
        noise = max(init_noise * (1. - decay_rate), 0.0)
        print('Noise std: %0.5f' % noise)
        model.noise_std = noise       # the captured train-step graph is keyed on this value (VanGan re-captures when it changes)

    def on_epoch_start(self, model, epoch, args, logs=None):
        """custom_callback.py:426-446."""
        self.set_learning_rate(model, epoch, args)
        self.updateDiscriminatorNoise(model.disc_I, model.layer_noise, epoch, args)
        self.updateDiscriminatorNoise(model.disc_S, model.layer_noise, epoch, args)

    def on_epoch_end(self, model, epoch, logs=None):
        """custom_callback.py:448-464."""
        a = self.imagePlotter(epoch, "genIS", self.imaging_val_data, self.imaging_val_full_vol_data, model.gen_IS, model.gen_SI,
                              process_img=True)
        b = self.imagePlotter(epoch, "geSI", self.segmentation_val_data, self.segmentation_val_full_vol_data, model.gen_SI,
                              model.gen_IS, outputFull=True)
        return a, b

    def imagePlotter(self, epoch, filename, setlist, dataset, genX, genY, nfig=6, outputFull=True, process_img=False, rng=None):
        """custom_callback.py:225-324.  The three forward passes of the panel (:265-267: prediction = genX(sample), cycled =
        genY(prediction), identity = genY(sample)) run on the CUDA path on a random crop of the first sample of `dataset`
        (an iterable of (volume, index) pairs); the four volumes are returned (and kept in `last_panels`).  The PNG is written
        when matplotlib is importable; the 3-D stitch of the full sample follows the reference's epoch rule (:321-324)."""
        sample, idx = next(iter(dataset))
        store = np.asarray(sample, dtype=np.float32)
        name = os.path.splitext(os.path.split(setlist[int(idx)])[1])[0] if setlist is not None else "sample"
        size = self.imgSize[1:5]
        rng = rng if rng is not None else np.random.default_rng()
        o = [int(rng.integers(0, store.shape[a] - size[a] + 1)) for a in range(3)]           # tf.image.random_crop
        crop = store[o[0]:o[0] + size[0], o[1]:o[1] + size[1], o[2]:o[2] + size[2], :][None]
        x = torch.as_tensor(crop, dtype=torch.float32, device=E.DEV).contiguous()
        if process_img and self.process_imaging_domain is not None:
            x = _apply_hook(self.process_imaging_domain, x, batched=True)
        prediction = genX(x, training=False)
        cycled = genY(prediction, training=False)
        identity = genY(x, training=False)
        panels = dict(sample=x[0].cpu().numpy(), prediction=prediction[0].cpu().numpy(), cycled=cycled[0].cpu().numpy(),
                      identity=identity[0].cpu().numpy(), name=name)
        self.last_panels = panels
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
            _, ax = plt.subplots(nfig + 1, 4, figsize=(12, 12))
            keys = ("sample", "prediction", "cycled", "identity")
            for j in range(nfig):
                for c, k in enumerate(keys):
                    ax[j, c].imshow(panels[k][:, :, j * int(size[2] / nfig), 0], cmap='gray')
                    ax[j, c].axis("off")
            for c, k in enumerate(keys):
                v = panels[k].ravel()
                ax[nfig, c].hist(v, bins=256, range=(float(v.min()), float(v.max())), fc='k', ec='k', density=True)
            os.makedirs("./GANMonitor", exist_ok=True)
            plt.savefig("./GANMonitor/{epoch}_{genID}.png".format(epoch=epoch + 1, genID=filename), dpi=300)
            plt.close()
        except ImportError:
            pass
        if epoch % self.period3D == 1 and outputFull and epoch > 160:
            self.stitch_subvolumes(genX, store, self.imgSize, epoch=epoch, name=name, process_img=process_img)
        return panels

    def stitch_subvolumes(self, gen, img, subvol_size, epoch=-1, stride=(25, 25, 128), name=None, output_path=None,
                          complete=False, padFactor=0.25, border_removal=True, process_img=False):
        """img: (H,W,D,1) float array (host).  gen: a generator model of this package.  Returns, on rank 0, the stitched prediction
        (float32, or uint8 when complete=False) exactly as the reference computes it before its TIFF write (None on other ranks).

        How the reference's serial loop (custom_callback.py:142-190) is executed: the window grid is enumerated exactly as there
        (including the repeated clamped window), every UNIQUE window is run through the generator once -- batched, forward-only,
        replayed from one captured CUDA graph -- and kept in HBM; then every output voxel adds the windows covering it in the
        reference's order (vg_stitch_gather_sum), so the result is bit-identical to the numpy loop driven by the same generator.
        Ranks take contiguous blocks of the unique-window list, upload only the rows of the volume their windows read, exchange the
        window outputs with one all-gather and each finalise a slab of rows; min / max are combined with a 2-float all-reduce."""
        import time
        import torch.distributed as dist
        prof = os.environ.get("VG_STITCH_PROFILE") == "1"
        marks = []

        def mark(label):
            if prof:
                torch.cuda.synchronize()
                marks.append((label, time.perf_counter()))
        mark("start")
        hook = self.process_imaging_domain if (process_img and self.process_imaging_domain is not None) else None
        img = np.asarray(img, dtype=np.float32)
        oshape = img.shape
        xs = ys = zs = 0
        if complete:                                                    # custom_callback.py:82-104
            # np.pad(img, ..., "symmetric") is never materialised: the windows are extracted from the un-padded volume with the
            # symmetric index map (vg_stitch_gather_sym); H, W, D below are the PADDED extents the enumeration runs over
            xs, ys = int(padFactor * img.shape[0]), int(padFactor * img.shape[1])
            if stride[2] != 1:
                zs = int(padFactor * img.shape[2])
        H0, W0, D0 = img.shape[0], img.shape[1], img.shape[2]
        assert img.shape[3] == 1
        H, W, D, Cc = H0 + 2 * xs, W0 + 2 * ys, D0 + 2 * zs, 1
        kH, kW, kD = subvol_size[1], subvol_size[2], subvol_size[3]
        if not complete or not border_removal:
            pH = pW = pD = 0
        else:
            pH, pW, pD = int(0.1 * kH), int(0.1 * kW), int(0.1 * kD)
            if kD == D:
                pD = 0
        sh, sw, sd = window_starts(H, kH, stride[0]), window_starts(W, kW, stride[1]), window_starts(D, kD, stride[2])
        starts = [(r, c, d) for r in sh for c in sw for d in sd]       # the reference's enumeration, duplicates included
        uniq, slot_of_start = [], {}
        for st in starts:
            if st not in slot_of_start:
                slot_of_start[st] = len(uniq)
                uniq.append(st)
        rank, world = self.strategy.rank, self.strategy.num_replicas_in_sync
        per = -(-len(uniq) // world)                                    # unique windows per rank (contiguous block, last one ragged)
        mine = uniq[rank * per:(rank + 1) * per]
        # slot of unique window u in the exchanged buffer: rank-major blocks of `per`
        slot_of = torch.tensor([slot_of_start[st] for st in starts], dtype=torch.int32, device=E.DEV)
        wins = torch.empty((per * world, kH, kW, kD), dtype=torch.float32, device=E.DEV)
        mark("host prep")
        if mine:
            lo, hi = min(st[0] for st in mine), max(st[0] for st in mine) + kH
            # only the rows this rank's windows read, uploaded in chunks on a copy stream while earlier windows run: a window batch
            # waits (event) for the last row it reads.  Padded (complete=True): rows of the un-padded volume, mapped symmetrically.
            sym = lambda p: -p - 1 if p < 0 else (2 * H0 - 1 - p if p >= H0 else p)
            if xs:
                src_rows = [sym(p - xs) for p in range(lo, hi)]
                lo0, hi0 = min(src_rows), max(src_rows) + 1
            else:
                lo0, hi0 = lo, hi
            src = torch.from_numpy(np.ascontiguousarray(img[lo0:hi0, :, :, 0]))
            vol = torch.empty((hi0 - lo0, W0, D0), dtype=torch.float32, device=E.DEV)
            copy_stream = self._copy_stream()
            copy_stream.wait_stream(torch.cuda.current_stream())
            up = {"rows": 0}

            def need_rows(p_hi):
                """enqueue the upload of every source row that padded rows [lo, p_hi) read and order the compute stream after it"""
                n = (max(sym(p - xs) for p in range(lo, min(p_hi, hi))) + 1 - lo0) if xs else (min(p_hi, hi) - lo)
                if n <= up["rows"]:
                    return
                with torch.cuda.stream(copy_stream):
                    vol[up["rows"]:n].copy_(src[up["rows"]:n], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                torch.cuda.current_stream().wait_event(ev)
                up["rows"] = n
            need_rows(lo + kH)
            mark("upload")
            B = min(self.window_batch, len(mine))
            fwd = None
            if len(mine) >= 2 * B and hook is None and self.use_graph:
                # captured once per (generator, window-batch shape): the graph reads the packed weights in place, so it stays
                # valid across optimizer steps and checkpoint loads
                key = (id(gen), B, kH, kW, kD)
                if key not in self._fwd_cache:
                    self._fwd_cache[key] = _GraphedForward(gen, (B, kH, kW, kD, 1))
                fwd = self._fwd_cache[key]
            mark("graph capture")
            for i in range(0, len(mine), B):
                chunk = mine[i:i + B]
                need_rows(max(a for a, _b, _c in chunk) + kH)
                padded = bool(xs or ys or zs)
                # un-padded: starts relative to the uploaded rows; padded: starts in padded coordinates (the kernel maps them)
                st = torch.tensor([((a if padded else a - lo), b_, c) for a, b_, c in chunk], dtype=torch.int32, device=E.DEV).reshape(-1)

                def gather(dst, nb):
                    if padded:
                        call("vg_stitch_gather_sym", vol, lo0, H0, W0, D0, xs, ys, zs, dst, st, nb, kH, kW, kD)
                    else:
                        call("vg_stitch_gather", vol, hi - lo, W, D, dst, st, nb, kH, kW, kD)
                if fwd is not None and len(chunk) == B:
                    gather(fwd.x, B)
                    out = fwd.replay()
                else:
                    win = torch.empty((len(chunk), kH, kW, kD, 1), dtype=torch.float32, device=E.DEV)
                    gather(win, len(chunk))
                    if hook is not None:                               # custom_callback.py:171-172, once per window
                        win = _apply_hook(hook, win, batched=False)
                    out = gen(win, training=False)
                wins[rank * per + i:rank * per + i + len(chunk)].copy_(out.reshape(len(chunk), kH, kW, kD))
            del vol
            mark("windows")
        if world > 1:
            dist.all_gather_into_tensor(wins, wins[rank * per:(rank + 1) * per].clone(), group=self.strategy.group)
        oH, oW, oD = (oshape[0], oshape[1], oshape[2]) if complete else (H, W, D)
        if complete and stride[2] == 1:
            oD = D
        rows_per = -(-oH // world)
        row0 = min(rank * rows_per, oH)
        rows = min(rows_per, oH - row0)
        slab = torch.empty((max(rows, 1), oW, oD), dtype=torch.float32, device=E.DEV)
        mm = torch.tensor([float("inf"), float("-inf")], dtype=torch.float32, device=E.DEV)
        enc = torch.empty(2, dtype=torch.int32, device=E.DEV)
        if rows > 0:
            axes = torch.tensor(list(sh) + list(sw) + list(sd), dtype=torch.int32, device=E.DEV)
            call("vg_stitch_gather_sum", wins, slot_of, axes, len(sh), len(sw), len(sd), kH, kW, kD, pH, pW, pD,
                 xs, ys, zs, row0, rows, oW, oD, slab, enc, 1)
            call("vg_stitch_minmax_decode", enc, mm)
        if world > 1:
            mn, mx = mm[0:1].clone(), mm[1:2].clone()
            dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=self.strategy.group)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.strategy.group)
            mm = torch.cat([mn, mx])
        as_u8 = not complete                                           # custom_callback.py:204-205 (astype('uint8'))
        if as_u8:
            res_d = torch.empty(slab.shape, dtype=torch.uint8, device=E.DEV)
            call("vg_stitch_scale_u8", slab, res_d, slab.numel(), mm)
        else:
            res_d = slab
            call("vg_stitch_scale", res_d, res_d.numel(), mm)
        if world > 1:                                                   # slabs -> rank 0
            full = torch.empty((rows_per * world, oW, oD), dtype=res_d.dtype, device=E.DEV) if rank == 0 else None
            padded = res_d if res_d.shape[0] == rows_per else torch.cat(
                [res_d[:rows], torch.zeros((rows_per - rows, oW, oD), dtype=res_d.dtype, device=E.DEV)])
            dist.gather(padded.contiguous(), list(full.split(rows_per)) if rank == 0 else None, dst=0, group=self.strategy.group)
            res_d = full[:oH] if rank == 0 else None
        self.last_stats = dict(windows=len(starts), unique=len(uniq), local_windows=len(mine))
        mark("finalize")
        if rank != 0:
            return None
        res = self._download(res_d)[..., None]
        mark("download")
        if prof:
            self.last_stats["phases_ms"] = {b[0]: round((b[1] - a[1]) * 1e3, 2) for a, b in zip(marks[:-1], marks[1:])}
            print("[stitch]", self.last_stats)
        if output_path is not None and name is not None:
            np.save(os.path.join(output_path, "{name}.npy".format(name=name)), res)
        return res

    def _copy_stream(self):
        if getattr(self, "_cstream", None) is None:
            self._cstream = torch.cuda.Stream()
        return self._cstream

    def _download(self, t):
        """device -> fresh numpy array through a cached pinned staging buffer, in chunks: the D2H copy of chunk i+1 (full PCIe rate)
        overlaps the host copy of chunk i into the result."""
        t = t.contiguous()
        out = np.empty(tuple(t.shape), dtype=np.uint8 if t.dtype == torch.uint8 else np.float32)
        flat_d, flat_h = t.view(-1), torch.from_numpy(out).view(-1)
        n, step = flat_d.numel(), (16 << 20) // t.element_size()
        if getattr(self, "_stage", None) is None or self._stage[0].dtype != t.dtype:
            self._stage = [torch.empty(step, dtype=t.dtype).pin_memory() for _ in range(2)]
        evs = [None, None]
        chunks = list(range(0, n, step))
        for k, off in enumerate(chunks + [None]):
            if off is not None:
                b = k & 1
                m = min(step, n - off)
                self._stage[b][:m].copy_(flat_d[off:off + m], non_blocking=True)
                evs[b] = torch.cuda.Event()
                evs[b].record()
            if k >= 1:
                pb, poff = (k - 1) & 1, chunks[k - 1]
                pm = min(step, n - poff)
                evs[pb].synchronize()
                flat_h[poff:poff + pm].copy_(self._stage[pb][:pm])
        return out

    def run_mapping(self, model, test_set, sub_img_size=(64, 64, 512, 1), segmentation=True, stride=(25, 25, 1),
                    padFactor=0.25, filetext=None, filepath=''):
        """custom_callback.py:466-509: every file of test_set (.npy volumes) through gen_IS (segmentation) or gen_SI."""
        results = []
        for imgdir in range(len(test_set)):
            img = np.load(test_set[imgdir])
            if img.ndim == 3:
                img = img[..., None]
            filename = os.path.splitext(os.path.basename(test_set[imgdir]))[0]
            gen = model.gen_IS if segmentation else model.gen_SI
            print(('Segmenting %s ... (%i / %i)' if segmentation else 'Mapping %s ... (%i / %i)') % (filename, imgdir + 1, len(test_set)))
            results.append(self.stitch_subvolumes(gen, img, sub_img_size, name=(filetext or "") + filename,
                                                  output_path=filepath or None, complete=True, process_img=not segmentation,
                                                  stride=stride, padFactor=padFactor))
        return results
