"""Secondary measurements of BASELINE.json configs 3, 4 and 5 on one GPU (config 5 also under torchrun).

  python scripts/bench_configs.py losses     # config 3: soft_skel / cycle-loss sweep, 64^3-256^3, iters 10-50
  python scripts/bench_configs.py vnet       # config 4: custom_vnet (gen_IS variant) forward+backward at 1x128^3
  python scripts/bench_configs.py sliding    # config 5: 512x512x256 volume, 128^3 windows, stride 64 (256 / 864 windows)

Every number is device time (CUDA events, median of 5 after 2 warm-ups, 256 MiB L2 flush between repetitions) and is printed
as one JSON line per case.  Roofline denominators come from MEASURED_PEAKS.json.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import Args, measured_peaks, synth_batch  # noqa: E402

TF_PEAK, HBM_PEAK, PEAK_SRC = measured_peaks()
_flush = None


def timed(fn, reps=5, warm=2):
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        _flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def run_losses():
    from van_gan_b200 import clDice_func as K, engine as E, loss_functions as LF

    class Cfg:
        global_batch_size, n_devices = 1, 1
        lambda_cycle, lambda_reconstruction, lambda_topology = 10.0, 5.0, 5.0
        loss_ctx = None

    rng = np.random.default_rng(4)
    for S in (64, 128, 256):
        x = torch.tensor(rng.random((1, S, S, S, 1)), dtype=torch.float32).cuda()
        g = torch.randn_like(x)
        V = S ** 3
        for iters in (10, 15, 25, 50):
            t_f = timed(lambda: K.soft_skel(x, iters))
            skel, bwd = K.soft_skel_with_grad(x, iters)
            t_b = timed(lambda: bwd(g))
            model = 16.0 * V * (iters + 1)          # one-iteration-per-pass byte model of the forward
            print(json.dumps({"config": 3, "case": "soft_skel", "S": S, "iters": iters, "fwd_ms": round(t_f, 4), "bwd_ms": round(t_b, 4),
                              "fwd_model_GBs": round(model / t_f / 1e6, 1), "fwd_frac_of_hbm": round(model / t_f / 1e6 / HBM_PEAK, 3),
                              "hbm_peak_GBs": HBM_PEAK}), flush=True)
        real = (x * 2 - 1).contiguous()
        cyc = torch.tanh(torch.randn_like(x))
        cfg = Cfg()

        def full_losses():
            cfg.loss_ctx = LF.LossContext()
            rv, cv = E.Var(real), E.Var(cyc)
            tot = (LF.cycle_seg_loss(cfg, rv, cv, iters=15) + LF.cycle_loss(cfg, rv, cv, typ="bce") + LF.cycle_loss(cfg, rv, cv, typ="mse")
                   + LF.cycle_reconstruction(cfg, rv, cv))
            return tot.seeds()                       # forward values + the gradient seeds w.r.t. the cycled volume (= backward)

        t = timed(full_losses)
        print(json.dumps({"config": 3, "case": "cycle_seg+bce+mse+ssim fwd+bwd (clDice iters 15)", "S": S, "ms": round(t, 3),
                          "Mvoxel_per_s": round(V / t / 1e3, 1)}), flush=True)


def run_vnet():
    from van_gan_b200 import engine as E
    from van_gan_b200.vnet_model import custom_vnet
    S = 128
    net = custom_vnet((S, S, S, 1), use_batch_norm=False, upsample_mode='upsample', dropout=0.5, filters=32, num_layers=4,
                      output_activation='tanh', seed=3)
    I, _ = synth_batch(1, S, 7)
    x = torch.tensor(I).cuda()
    gy = torch.randn((1, S, S, S, 1), device="cuda")

    def step():
        tape = E.Tape()
        out = net.forward(tape, E.Var(x), training=True, seed=1)
        net.zero_grad()
        tape.backward([(out, gy)], net.trainable_variables)

    def fwd():
        net.forward(E.Tape(enabled=False), E.Var(x), training=True, seed=1)

    # the eager step is ~700 launches enqueued from Python: on a busy host it is enqueue-bound (12.1 / 12.6 / 17.5 / 82 ms were measured
    # for the same kernels in different processes); inside VanGan.train_step the same network runs from the captured graph
    t_f, t = timed(fwd), timed(step, reps=9)
    fl_f = 1369.96e9                                # forward FLOPs of the gen_IS variant at 128^3
    print(json.dumps({"config": 4, "case": "custom_vnet gen_IS (IN, upsample, f=32) 1x128^3 bf16", "fwd_ms": round(t_f, 3),
                      "fwd_bwd_ms": round(t, 3), "launch": "eager, enqueued from Python (median of 9)", "fwd_TFLOPs": round(fl_f / t_f / 1e9, 1),
                      "fwd_bwd_TFLOPs": round(3 * fl_f / t / 1e9, 1),
                      "frac_of_bf16_peak_fwd": round(fl_f / t_f / 1e9 / TF_PEAK, 3), "bf16_peak_TFLOPs": TF_PEAK}), flush=True)


def sliding_volume(shape=(512, 512, 256)):
    """Photoacoustic-like synthetic volume: the 128^3 recipe of bench.synth_batch tiled (mirror-free) to the full extent."""
    I, _ = synth_batch(1, 128, 5)
    reps = [(s + 127) // 128 for s in shape]
    return np.tile(I[0, ..., 0], reps)[:shape[0], :shape[1], :shape[2], None].astype(np.float32)


def run_sliding(gen=None, strategy=None, cases=((False, "stride 64, complete=False"), (True, "stride 64, complete=True padFactor 0.25")),
                window_batch=8):
    """Returns the list of result dicts (rank 0 prints them).  Wall time of the public call: host volume in, host result out."""
    import time
    from van_gan_b200.custom_callback import GanMonitor
    from van_gan_b200.distribute import Strategy
    from van_gan_b200.resunet_model import ResUNet
    strategy = strategy or Strategy()
    if gen is None:
        gen = ResUNet((128, 128, 128, 1), upsample_mode='simple', dropout_type='none', seed=1234)
    mon = GanMonitor(strategy=strategy, window_batch=window_batch)
    vol = sliding_volume()
    out = []
    for complete, label in cases:
        best = None
        for rep in range(2):                         # first call warms the allocator / kernels
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            mon.stitch_subvolumes(gen, vol, (1, 128, 128, 128, 1), stride=(64, 64, 64), complete=complete, padFactor=0.25)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        st = mon.last_stats
        out.append({"config": 5, "case": label, "windows": st["windows"], "unique": st["unique"], "n_gpus": strategy.num_replicas_in_sync,
                    "seconds": round(best, 4), "Mvoxel_per_s": round(vol.size / best / 1e6, 1),
                    "window_Mvoxel_per_s": round(st["windows"] * 128 ** 3 / best / 1e6, 1),
                    "gen_fwd_TFLOPs": round(st["windows"] * 299.31e9 / best / 1e12, 1)})
        if "phases_ms" in st:                        # VG_STITCH_PROFILE=1 (synchronising marks: the total above is then not a bench value)
            out[-1]["phases_ms"] = st["phases_ms"]
    return out


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if which in ("losses", "all"):
        run_losses()
    if which in ("vnet", "all"):
        run_vnet()
    if which in ("sliding", "all"):
        from van_gan_b200.distribute import Strategy, init_from_env
        rank, world, local = init_from_env()
        for r in run_sliding(strategy=Strategy()):
            if rank == 0:
                print(json.dumps(r), flush=True)
