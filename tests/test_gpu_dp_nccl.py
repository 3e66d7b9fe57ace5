"""N = 2 data-parallel parity of the CUDA path over NCCL against the oracle's MirroredStrategy emulation
(oracle.step.train_step_dp: shard the global batch, per-replica losses with the reference's replica scalings, SUM the
gradients and the result dict, then one clip+Adam per network -- vangan.py:426-438,459-490).

Needs two GPUs (skipped otherwise; the round's 2-GPU run is kept under profiles/).  One process per GPU (mp.spawn),
`VanGan.distributed_train_step` on each rank's shard with explicit per-replica discriminator noise / dropout tensors.
Checks on rank 0: the ten SUM-reduced losses at 2e-2, the all-reduced gradients against the emulated and the fp32 oracle
(same criteria as tests/test_gpu_parity_r2.py), and that both ranks hold bit-identical weights after the update.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, S, b, outdir):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from test_gpu_parity_r2 import step_criteria
    from oracle import losses as OL, nets as ON, step as OS
    from test_gpu_train_step import Args, _setup
    from van_gan_b200.distribute import Strategy, init_from_env
    from van_gan_b200.vangan import VanGan
    init_from_env("nccl")
    G = b * world
    real_I, real_S, init, _ = _setup(S, G, world, 300)
    rng = np.random.default_rng(301)
    rands = [{k: ON.make_disc_rand(rng, b, S) for k in ("S_real", "S_fake", "I_real", "I_fake")} for _ in range(world)]
    gan = VanGan(Args(S, G, world), strategy=Strategy(), gen_i2s='resUnet', gen_s2i='resUnet')
    assert gan.strategy.num_replicas_in_sync == world
    for k, net in gan.networks.items():
        net.load(init[k])
    sl = slice(rank * b, (rank + 1) * b)
    rand_d = {k: ([t.cuda() for t in nz], [m.cuda() for m in mk]) for k, (nz, mk) in rands[rank].items()}
    res_k = gan.distributed_train_step(real_I[sl], real_S[sl], rand=rand_d)
    # net.g holds the all-reduced (summed) gradients; weights are post-Adam
    sums = torch.stack([net.w.double().sum() for net in gan.networks.values()] + [net.g.double().sum() for net in gan.networks.values()])
    gathered = [torch.zeros_like(sums) for _ in range(world)]
    dist.all_gather(gathered, sums)
    assert all(torch.equal(gathered[0], g) for g in gathered), "replicas diverged"
    if rank == 0:
        cfg = OL.make_cfg(G, world)

        def oracle(emu):
            ON.Emu.on = emu
            try:
                P = {k: ON.to_torch(v) for k, v in init.items()}
                opts = {k: OS.Adam(list(v.keys())) for k, v in init.items()}
                res, g = OS.train_step_dp(cfg, P, opts, real_I, real_S, rands)
                return res, g, P
            finally:
                ON.Emu.on = False
        res_o, g_o, P_o = oracle(False)
        _, g_e, _ = oracle(True)
        lines = []
        for k in OS.RESULT_KEYS:
            assert abs(res_k[k] - res_o[k]) <= 2e-2 * abs(res_o[k]) + 1e-4, (k, res_k[k], res_o[k])
        for name, net in gan.networks.items():
            ok, e = step_criteria(name, net.export_grads(), g_o[name], g_e[name])
            lines.append("N=2 %-7s: cos(CUDA, fp32) %.3f | norm ratio %.3f | CUDA vs fp32 %.3f | Emu vs fp32 %.3f | CUDA vs Emu %.3f"
                         % (name, e["cos_fp32"], e["norm_ratio"], e["vs_fp32"], e["emu_vs_fp32"], e["vs_emu"]))
            assert ok, (name, e)
            # the update itself: displacement correlates with the oracle's (Adam's first step is lr*sign(g))
            w = net.export()
            num = den1 = den2 = 0.0
            for n in w:
                dk = torch.tensor(w[n] - init[name][n]).double().flatten()
                do = (P_o[name][n].detach() - torch.tensor(init[name][n])).double().flatten()
                num += float(dk @ do); den1 += float(dk @ dk); den2 += float(do @ do)
            lines.append("N=2 %-7s: weight-displacement cosine vs oracle %.3f" % (name, num / (den1 * den2) ** 0.5))
            assert num / (den1 * den2) ** 0.5 > 0.5
        with open(os.path.join(outdir, "dp_nccl.txt"), "w") as f:
            f.write("\n".join(lines) + "\n")
        print("\n".join(lines))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_two_ranks_nccl_matches_oracle(cuda, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), 64, 1, str(tmp_path)), nprocs=2, join=True)
    print(open(tmp_path / "dp_nccl.txt").read())


def _worker_noise(rank, world, port, outdir):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from test_gpu_train_step import Args, synth
    from van_gan_b200.distribute import Strategy, init_from_env
    from van_gan_b200.vangan import VanGan
    init_from_env("nccl")
    S, b = 32, 1
    gan = VanGan(Args(S, b * world, world), strategy=Strategy(), gen_i2s='resUnet', gen_s2i='resUnet')
    rng = np.random.default_rng(7)
    I, Sg = synth(rng, 1, S)           # the SAME sample on both ranks: only the noise / dropout draws can differ
    losses = []
    for _ in range(4):                 # two eager steps, capture, replays: the graph path with its in-graph all-reduce
        gan.train_step(I.cuda(), Sg.cuda())
    off = torch.tensor([int(gan._seed_dev.item())], dtype=torch.int64, device="cuda")
    offs = [torch.zeros_like(off) for _ in range(world)]
    dist.all_gather(offs, off)
    assert len({int(o.item()) for o in offs}) == world, "replicas share one noise / dropout stream"
    sums = torch.stack([net.w.double().sum() for net in gan.networks.values()])
    gathered = [torch.zeros_like(sums) for _ in range(world)]
    dist.all_gather(gathered, sums)
    assert all(torch.equal(gathered[0], g) for g in gathered), "replicas diverged on the graph path"
    dist.barrier()
    dist.destroy_process_group()


def test_dp_replicas_draw_independent_noise_and_stay_in_sync(cuda, tmp_path):
    """MirroredStrategy draws GaussianNoise / SpatialDropout3D independently per replica; weights stay identical."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker_noise, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
