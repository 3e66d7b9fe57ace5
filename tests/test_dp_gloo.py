"""CPU, world_size 2, gloo: the data-parallel plumbing (van_gan_b200.distribute.Strategy) sums per-replica
gradients and result dicts exactly like the single-process shard-and-sum emulation of MirroredStrategy."""
import os
import socket
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    from van_gan_b200.distribute import Strategy, init_from_env
    from oracle import losses as OL, nets as ON, step as OS
    r, w, _ = init_from_env(backend="gloo")
    strat = Strategy()
    assert strat.num_replicas_in_sync == world and strat.rank == rank
    data = _make_case()
    cfg = OL.make_cfg(world, world)
    P = {k: ON.to_torch(v) for k, v in data["init"].items()}
    sl = slice(rank, rank + 1)
    res, grads, _ = OS.replica_grads(cfg, P, data["I"][sl], data["S"][sl], data["rands"][rank], iters=3)
    flat = torch.cat([g.reshape(-1) for g in grads["gen_IS"].values()])
    h = strat.all_reduce_async(flat)           # what VanGan.train_step launches per network
    h.wait()
    vals = torch.tensor([float(v.detach()) for v in res.values()], dtype=torch.float64)
    strat.reduce("SUM", vals)                  # VanGan.reduce_dict
    if rank == 0:
        torch.save({"flat": flat, "vals": vals}, out)
    torch.distributed.destroy_process_group()


def _make_case():
    from oracle import nets as ON
    rng = np.random.default_rng(11)
    S, G = 32, 2
    init = {"gen_IS": ON.init_params(ON.resunet_param_shapes(), 1), "gen_SI": ON.init_params(ON.resunet_param_shapes(), 2),
            "disc_I": ON.init_params(ON.disc_param_shapes(), 3), "disc_S": ON.init_params(ON.disc_param_shapes(), 4)}
    I = torch.tensor(rng.random((G, S, S, S, 1)) * 2 - 1, dtype=torch.float32)
    Sg = torch.tensor(rng.random((G, S, S, S, 1)) * 2 - 1, dtype=torch.float32)
    rands = [{k: ON.make_disc_rand(rng, 1, S) for k in ("S_real", "S_fake", "I_real", "I_fake")} for _ in range(G)]
    return dict(init=init, I=I, S=Sg, rands=rands)


def test_two_replicas_allreduce_matches_shard_and_sum(tmp_path):
    from oracle import losses as OL, nets as ON, step as OS
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    torch.set_num_threads(2)      # same reduction order as the workers: even the fp32 oracle is chaotic at the 1e-3 level
    data = _make_case()
    cfg = OL.make_cfg(2, 2)
    P = {k: ON.to_torch(v) for k, v in data["init"].items()}
    res, grads = OS.train_step_dp(cfg, P, None, data["I"], data["S"], data["rands"], iters=3)
    flat = torch.cat([g.reshape(-1) for g in grads["gen_IS"].values()])
    rel = float((got["flat"] - flat).norm() / flat.norm())
    assert rel < 1e-3, rel
    assert np.allclose(got["vals"].numpy(), np.array(list(res.values())), rtol=1e-6)
