"""CPU checks of the teacher-forced block decomposition used by the GPU parity tests (tests/_blocks.py): chaining the
blocks reproduces the oracle's whole-network forward bit for bit, the parameter lists partition the networks, and the
random-projection estimator used by the 128^3 golden test recovers a known relative error."""
import numpy as np
import torch

import _blocks as B
from oracle import nets as ON


def test_generator_blocks_chain_to_the_network():
    rng = np.random.default_rng(0)
    P = ON.to_torch(ON.init_params(ON.resunet_param_shapes(), 1, 0.05), requires_grad=False)
    x = torch.tensor(rng.standard_normal((1, 32, 32, 32, 1)), dtype=torch.float32)
    taps = {"input": x}
    for name, (fn, ins, _) in B.gen_blocks().items():
        taps[name] = fn(P, *[taps[i] for i in ins])
    assert torch.equal(taps["head"], ON.resunet_forward(P, x))
    owned = [n for _, (_, _, pre) in B.gen_blocks().items() for n in P if any(n.startswith(q) for q in pre)]
    assert sorted(owned) == sorted(P.keys())


def test_discriminator_stages_chain_to_the_network():
    rng = np.random.default_rng(1)
    P = ON.to_torch(ON.init_params(ON.disc_param_shapes(), 3, 0.05), requires_grad=False)
    x = torch.tensor(rng.standard_normal((2, 16, 16, 16, 1)), dtype=torch.float32)
    nz, mk = ON.make_disc_rand(rng, 2, 16)
    ins = B.disc_stage_inputs(P, x, nz, mk)
    y = B.disc_stage(4)(P, ins[4], nz, mk)
    assert torch.allclose(y, ON.disc_forward(P, x, nz, mk), rtol=0, atol=0)
    owned = [n for k in range(5) for n in B.disc_stage_params(k)]
    assert sorted(owned) == sorted(P.keys())


def test_projection_estimator():
    rng = np.random.default_rng(2)
    a = torch.tensor(rng.standard_normal(200000))
    b = a + 0.3 * torch.tensor(rng.standard_normal(200000))
    pa, pb = B.projections(a, 5), B.projections(b, 5)
    rel = float(np.sqrt(np.mean((pb - pa) ** 2))) / float(a.norm())
    cs = float(np.mean(pa * pb)) / float(a.norm() * b.norm())
    true_rel = float((b - a).norm() / a.norm())
    true_cs = float(a @ b / (a.norm() * b.norm()))
    assert abs(rel - true_rel) < 0.25 * true_rel and abs(cs - true_cs) < 0.2
    assert np.array_equal(pa, B.projections(a, 5))


def test_seed_offsets_are_unique_per_step_and_replica():
    import importlib
    import sys
    # vangan imports the engine (torch only at import time; no CUDA call until a network is built)
    from van_gan_b200.vangan import VanGan
    seen = set()
    for step in range(50):
        for rank in range(8):
            off = VanGan.seed_offset(1234, step, 8, rank)
            assert off % 64 == 0 and off not in seen
            seen.add(off)
    assert VanGan.seed_offset(1234, 3, 1, 0) == (1234 * 1000003 + 3) * 128
