#!/usr/bin/env python
"""bench.py — VAN-GAN 128^3 train-step throughput (volumes/s) on N B200s of one node.

Workload (BASELINE.json configs[1]): full VanGan.train_step — 2 ResUNet generators + 2 3D-PatchGAN
discriminators, all ten losses (LSGAN, BCE/MSE cycle, SSIM reconstruction, clDice iters=15), four backward
sweeps, gradient all-reduce, clip+Adam — on synthetic 128^3 single-channel volumes, GLOBAL batch 8, sharded
b = 8/N per GPU (strong scaling), one process per GPU over NCCL.

  python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # the CPU oracle port of the reference, host cores
  torchrun ... bench.py --gpus N ...                       # N > 1 (driver-launched)

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train_step_volumes_per_s_128cubed"
UNIT = "volumes/s"


class Args:
    """The attributes VanGan reads from the reference's `args` namespace (main.py:62-105)."""

    def __init__(self, S, G, nd):
        self.N_DEVICES, self.GLOBAL_BATCH_SIZE = nd, G
        self.INPUT_IMG_SIZE = (G, S, S, S, 1)
        self.CHANNELS, self.DIMENSIONS = 1, 3
        self.SUBVOL_PATCH_SIZE = (S, S, S)
        self.train_steps, self.BATCH_SIZE, self.output_dir = 1, G // nd, "/tmp"


def synth_batch(n, S, seed):
    """Synthetic photoacoustic-like imaging volumes and vessel-like segmentation volumes in [-1,1]
    (SURVEY.md 8d config 2): smoothed noise; union of random soft-edged tubes."""
    import numpy as np
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    I = np.empty((n, S, S, S, 1), np.float32)
    Sg = np.empty((n, S, S, S, 1), np.float32)
    ax = np.arange(S, dtype=np.float32)
    zz, yy, xx = np.meshgrid(ax, ax, ax, indexing="ij")
    for i in range(n):
        v = ndimage.gaussian_filter(rng.standard_normal((S, S, S)).astype(np.float32), 2.0)
        I[i, ..., 0] = 2 * (v - v.min()) / (v.max() - v.min()) - 1
        t = np.zeros((S, S, S), np.float32)
        for _ in range(24):
            p0, d = rng.random(3) * S, rng.standard_normal(3)
            d /= np.linalg.norm(d)
            rz, ry, rx = zz - p0[0], yy - p0[1], xx - p0[2]
            along = rz * d[0] + ry * d[1] + rx * d[2]
            dist = np.sqrt(np.maximum(rz * rz + ry * ry + rx * rx - along * along, 0))
            t = np.maximum(t, 1.0 / (1.0 + np.exp((dist - (1.5 + 2.5 * rng.random())) * 2.0)))
        Sg[i, ..., 0] = 2 * t - 1 + 1e-3 * rng.standard_normal((S, S, S)).astype(np.float32)
    return I, Sg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nme, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nme)
            except (ValueError, IndexError):
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bench_config(S, G, world):
    """The `config` object both arms print (the reference arm runs on OUR arm's config)."""
    return {"workload": "VanGan.train_step 2xResUNet(f16,L4)+2xPatchGAN(f64), %d^3x1 volumes, global batch %d "
                        "(b=%d per GPU), clDice iters 15, LSGAN, Adam+clipnorm" % (S, G, G // world),
            "parallelism": "dp%d" % world, "l2": "256 MiB flush write between timed steps"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1375.8), d.get("hbm_gbs", 6551.7), "measured"
    return 1400.0, 6650.0, "fallback"


# ------------------------------------------------------------------------------------------ CPU arms
def _host_threads():
    """All host threads, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to every rank)."""
    import torch
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0)) or n
    except (AttributeError, OSError):
        pass
    torch.set_num_threads(n)
    return torch.get_num_threads()


def _oracle_state(seed0=1234):
    from oracle import nets as ON, step as OS
    P = {"gen_IS": ON.to_torch(ON.init_params(ON.resunet_param_shapes(), seed0)),
         "gen_SI": ON.to_torch(ON.init_params(ON.resunet_param_shapes(), seed0 + 1)),
         "disc_I": ON.to_torch(ON.init_params(ON.disc_param_shapes(), seed0 + 2)),
         "disc_S": ON.to_torch(ON.init_params(ON.disc_param_shapes(), seed0 + 3))}
    return P, {k: OS.Adam(list(v.keys())) for k, v in P.items()}


def cpu_oracle_steps(S, steps, warmup=0):
    """Times `steps` train steps of the oracle port of the reference (torch CPU fp32, all host threads) on one 1 x S^3 sample.
    Returns (mean seconds per step, threads)."""
    import numpy as np
    import torch
    from oracle import losses as OL, nets as ON, step as OS
    cores = _host_threads()
    rng = np.random.default_rng(0)
    I, Sg = synth_batch(1, S, 11)
    real_I, real_S = torch.tensor(I), torch.tensor(Sg)
    P, opts = _oracle_state()
    cfg = OL.make_cfg(1, 1)
    times = []
    for it in range(warmup + steps):
        rand = {k: ON.make_disc_rand(rng, 1, S) for k in ("S_real", "S_fake", "I_real", "I_fake")}
        t0 = time.perf_counter()
        OS.train_step_dp(cfg, P, opts, real_I, real_S, [rand])
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times), cores


def cpu_full_size_ok():
    """One oracle step at 1 x 128^3 keeps ~45 GB of autograd state: only run it where the host has the memory."""
    try:
        import psutil
        return psutil.virtual_memory().available > 70 * 2 ** 30
    except Exception:
        return False


def cpu_baseline_128(allow_full=True):
    """CPU baseline on the bench's own volume size.  Preferred: ONE full oracle train_step on 1 x 128^3 (the workload processes 8
    such volumes per step; CPU time is linear in the batch).  Fallback when the host lacks the memory: a 1 x 64^3 step scaled
    by the voxel ratio 8.  Returns (volumes/s, threads, description, seconds of the sample)."""
    if allow_full and cpu_full_size_ok():
        sec, cores = cpu_oracle_steps(128, 1)
        return 1.0 / sec, cores, "one full oracle (torch CPU fp32) train_step on 1x128^3: %.1f s" % sec, sec
    sec, cores = cpu_oracle_steps(64, 1)
    return 1.0 / (8.0 * sec), cores, ("one oracle (torch CPU fp32) train_step on 1x64^3 (%.1f s) scaled by the voxel ratio 8 "
                                      "(host memory < 70 GB free: the 1x128^3 step does not fit)" % sec), sec


def run_reference(args):
    """The reference arm: the reference's own CPU path for this workload.  TensorFlow cannot be installed in this image, so it is
    the oracle port (kind "port").  Timed steps are bounded samples (one 1 x 64^3 train step each) so that --steps 20 --warmup 5
    ends within minutes; the VALUE comes from one full 1 x 128^3 step measured in the same run whenever the host has the memory
    for it, so the line is on our arm's configuration (128^3 volumes), not an extrapolation."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = _host_threads()
    v128, _, desc128, sec128 = cpu_baseline_128()
    sec64, _ = cpu_oracle_steps(64, args.steps, max(args.warmup - 1, 0))
    v64 = 1.0 / (8.0 * sec64)
    full = "1x128^3" in desc128 and "scaled" not in desc128
    sample = "%s -> %.4f volumes/s (the line's value); the %d timed steps are bounded samples: one oracle train_step on 1x64^3 each " \
             "(%.2f s/step; x8 voxels -> %.4f volumes/s)" % (desc128, v128, args.steps, sec64, v64)
    line = {"impl": "reference", "metric": METRIC, "value": v128, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec64 * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args.size, args.global_batch, max(world, args.gpus)),
            "note": "TensorFlow is not installable here; the reference arm is the oracle port (torch CPU fp32) of the reference's "
                    "train_step on the host cores; ms_per_step = mean duration of a timed step (= one bounded 1x64^3 sample, 1/8 volume)",
            "full_128_step_s": sec128 if full else None, "value_from_64cubed_samples": v64,
            "cpu_baseline": {"value": v128, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v128, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from van_gan_b200 import _lib
    from van_gan_b200.distribute import Strategy, init_from_env
    from van_gan_b200.vangan import VanGan

    rank, world, local = init_from_env()
    torch.cuda.set_device(local)
    _lib.lib()
    G, S = args.global_batch, args.size
    assert G % world == 0, "global batch must divide over the ranks"
    b = G // world
    strategy = Strategy()
    gan = VanGan(Args(S, G, world), strategy, gen_i2s='resUnet', gen_s2i='resUnet', seed=1234)

    # host-resident (pinned) shard of the synthetic global batch
    I, Sg = synth_batch(b, S, 100 + rank)
    hI, hS = torch.from_numpy(I).pin_memory(), torch.from_numpy(Sg).pin_memory()
    dI, dS = hI.cuda(), hS.cuda()
    l2_flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, prof=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _lib.PROFILER = prof
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        _lib.PROFILER = None
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_resident():
        l2_flush.zero_()          # flush L2 between timed iterations (inputs are << L2 at b=1)
        gan.distributed_train_step(dI, dS)

    def step_e2e():
        l2_flush.zero_()
        gan.distributed_train_step(hI, hS)    # H2D of the shard inside; the result dict is read back (D2H)

    for _ in range(args.warmup):
        step_resident()            # the third call captures the whole step into a CUDA graph; later calls replay it
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = _lib.lib().vg_launch_count()
    ms = timed(step_resident, args.steps)
    launches = _lib.lib().vg_launch_count() - l0
    graph_on = gan._graph is not None
    launch_mode = "eager launches"
    if graph_on:
        launch_mode = ("CUDA graph replay: one graph (losses, four backward sweeps, clip+Adam, operand repack)" if world == 1 else
                       ("CUDA graph replay: one graph with the bucketed NCCL all-reduces (vg_comm) captured on the communication stream"
                        if gan._graph["mode"] == "single" else
                        "CUDA graph replay: two graphs (forward + the four backward sweeps side by side / clip+Adam); the bucketed NCCL "
                        "all-reduces (vg_comm) of the four networks are enqueued on the communication stream between the replays"
                        if gan._graph["mode"] == "two" else
                        "CUDA graph replay: four graphs (forward + generator sweeps / discriminator sweeps / clip+Adam of the generators / "
                        "clip+Adam of the discriminators); the bucketed NCCL all-reduces (vg_comm) are enqueued on the communication stream "
                        "between the replays: the generators' run beside the discriminator sweeps, the discriminators' beside the generators' "
                        "update"))
        if gan.use_streams:
            launch_mode += "; the two generator chains of the forward pass and the backward sweeps run on side streams"
    comm_msgs = int(strategy._L.vg_comm_collectives(strategy.comm)) if strategy.comm is not None else 0
    if graph_on:                   # replays do not pass through the host-side launch counter: count what the graph holds
        launches = gan.launches_per_replay * args.steps
    clk = clocks.stop() if rank == 0 else None
    ms_e2e = timed(step_e2e, args.steps)
    # per-kernel-family device times: a separate EAGER pass (CUDA events around every ABI call cannot be recorded inside a
    # graph replay); same inputs, same kernels, same launch order
    # ... and on ONE stream: with the branches of the step on side streams a kernel's event-to-event time would include the kernels
    # of other branches that share the SMs with it
    saved = (gan._graph, gan.use_graph, gan._side, gan._wg_side)
    gan._graph, gan.use_graph, gan._side, gan._wg_side = None, False, None, None
    psteps = min(args.steps, 2)
    step_resident()
    prof = _lib.Profiler([], detail=True)   # keyed by (call, shape)
    ms_prof = timed(step_resident, psteps, prof)
    fam = prof.summary()
    gan._graph, gan.use_graph, gan._side, gan._wg_side = saved

    # second half of BASELINE.json's metric: sliding-window inference (config 5) through the public GanMonitor call, host volume
    # in / host result out, windows sharded over the ranks
    sliding = None
    if not args.no_sliding and S == 128:
        import importlib.util
        spec = importlib.util.spec_from_file_location("vg_bench_configs", os.path.join(ROOT, "scripts", "bench_configs.py"))
        bc = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bc)
        torch.cuda.empty_cache()
        sliding = bc.run_sliding(gen=gan.gen_IS, strategy=strategy, cases=((False, "512x512x256, 128^3 windows, stride 64, complete=False"),))[0]
    if world > 1:
        dist.barrier()
        strategy.destroy()
        dist.destroy_process_group()
    if rank != 0:
        return
    tf_peak, hbm_peak, peak_src = measured_peaks()
    value = G * args.steps / (ms / 1e3)
    e2e = G * args.steps / (ms_e2e / 1e3)
    # family totals and the dominant kernel: tc_conv_kernel (tcgen05 implicit GEMM) runs every stride-1 forward with
    # Cin,Cout % 16 == 0 and every dgrad with Cin,Cout % 16 == 0 (one launch per call, 8 per stride-2 dgrad)
    import re
    families, tc_ms, tc_flop, tc_calls, conv_ms = {}, 0.0, 0.0, 0, 0.0
    for key, v in fam.items():
        name = key.split(" ")[0]
        families[name] = families.get(name, 0.0) + v["ms"]
        m = re.match(r"vg_conv3d_(fwd|dgrad) (\d+)->(\d+) k(\d)s(\d)", key)
        if name.startswith("vg_conv3d"):
            conv_ms += v["ms"]
        if m:
            kind, ci, co, kk, st = m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5))
            k1_small = (ci, co) in ((48, 16), (96, 32), (16, 16), (32, 32))     # 1x1x1 forward shapes served by conv_small.cu
            # tcgen05 path: every k3 / k4 layer; 1x1x1: stride-2 input gradients and the stride-1 shapes without a streaming kernel
            on_tc = kk >= 3 or (kk == 1 and not (st == 1 and k1_small) and (kind == "dgrad" or st == 1))
            if ci % 16 == 0 and co % 16 == 0 and on_tc:
                tc_ms += v["ms"]; tc_flop += v["work"]; tc_calls += v["calls"]
    achieved = tc_flop / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else None

    def fam_rate(prefix, scale):
        t = sum(v["ms"] for k, v in fam.items() if k.startswith(prefix))
        w = sum(v["work"] for k, v in fam.items() if k.startswith(prefix))
        return (w / (t / 1e3) / scale, t / psteps) if t > 0 and w > 0 else (None, t / psteps)

    extra = {}
    for label, prefix, bound in (("instnorm_bwd", "vg_instnorm_bwd", "hbm"), ("instnorm_apply", "vg_instnorm_apply", "hbm"),
                                 ("instnorm_stats", "vg_instnorm_stats", "hbm"), ("soft_skel_fwd", "vg_soft_skel_fwd", "hbm"),
                                 ("soft_skel_bwd", "vg_soft_skel_bwd", "hbm"), ("conv3d_wgrad", "vg_conv3d_wgrad", "tensor")):
        rate, t = fam_rate(prefix, 1e9 if bound == "hbm" else 1e12)
        pk = hbm_peak if bound == "hbm" else tf_peak
        extra[label] = {"bound": bound, "achieved": rate, "peak": pk, "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
                        "frac": (rate / pk) if rate else None, "ms_per_step": round(t, 3)}
    ms_ref = ms_prof / psteps      # step time of the profiled (eager) pass: shares are taken against it
    roofline = {"kernel": "tc_conv_kernel (tcgen05/TMEM implicit-GEMM Conv3D: stride-1/2 forward + all dgrads)", "bound": "tensor",
                "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": (achieved / tf_peak) if achieved else None,
                # one `ncu --set full` capture of this kernel (d-march, BD=8) on the 48->16 k3 layer at 8x128^3
                # (profiles/r02_ncu_full_fwd_48-16_call23.txt): dram read 1.6875 GB + write 0.5174 GB for that launch; its
                # algorithmic bytes (bf16 in + out) are 2.22e9
                "traffic": 2.2049e9, "traffic_launch": "fwd 48->16 k3 s1, 8x130^3 -> 8x128^3",
                "traffic_source": "ncu --set full capture of that one launch, profiles/r02_ncu_full_fwd_48-16_call23.txt (not re-measured by this run)",
                "peak_source": "%s bf16 sustained (kernel timed inside a long step)" % peak_src,
                "calls_per_step": tc_calls / psteps, "share_of_step": tc_ms / psteps / ms_ref if ms_ref > 0 else None,
                "conv_family_share_of_step": conv_ms / psteps / ms_ref if ms_ref > 0 else None,
                "eager_profiled_ms_per_step": ms_ref,
                "families_ms_per_step": {k: round(v / psteps, 3) for k, v in sorted(families.items(), key=lambda kv: -kv[1])[:14]}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": bench_config(S, G, world),
            "launch_mode": launch_mode,
            "clocks": clk, "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(2 * b * S ** 3 * 4 * world),
                                   "d2h_bytes_per_step": int(64 * 8 * world), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "roofline": roofline, "roofline_other_kernels": extra,
            "peak_mem_gib": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}
    if world > 1:
        line["comm"] = {"library": "vg_comm (NCCL %d resolved at run time)" % _lib.lib().vg_comm_nccl_version() if strategy.use_vg_comm else "torch.distributed",
                        "messages_enqueued_outside_graph_replays": comm_msgs, "bucket_elems": __import__("van_gan_b200.distribute", fromlist=["x"]).BUCKET_ELEMS}
    if sliding is not None:
        line["sliding_window"] = {"value": sliding["Mvoxel_per_s"], "unit": "Mvoxel/s", "windows": sliding["windows"],
                                  "seconds": sliding["seconds"], "gen_fwd_TFLOPs": sliding["gen_fwd_TFLOPs"], "workload": sliding["case"]}
    if world == 1 and not args.no_cpu_baseline:
        torch.cuda.empty_cache()
        v, cores, desc, _sec = cpu_baseline_128(allow_full=S == 128)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--global-batch", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sliding", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
