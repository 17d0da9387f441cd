"""world_size-2 gloo test of the sharded objective (the N>1 path of bench.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from conftest import ROOT, load_model


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fakelib import FakeLib

        from metran_b200.batched import KalmanEngine
        from metran_b200.distributed import gather_per_model, pooled_objective, shard_bounds

        g = load_model("dfm_notebook")
        eng = KalmanEngine(lib=FakeLib(), device="cpu")
        draws = np.concatenate([g["alpha_init"][None], g["alpha_opt"][None], g["alpha_draws"][:5]])
        lo, hi = shard_bounds(len(draws), rank, world)
        y = g["y"][:400]
        mle, _, _ = eng.dfm_loglik_alpha(y, draws[lo:hi], g["factors"])
        pooled = pooled_objective(eng, mle)
        full = gather_per_model(mle, len(draws))
        ret[rank] = (float(pooled[0]), full.numpy().copy(), (lo, hi))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_objective_matches_single_process():
    from fakelib import FakeLib

    from metran_b200.batched import KalmanEngine

    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    g = load_model("dfm_notebook")
    draws = np.concatenate([g["alpha_init"][None], g["alpha_opt"][None], g["alpha_draws"][:5]])
    eng = KalmanEngine(lib=FakeLib(), device="cpu")
    mle, _, _ = eng.dfm_loglik_alpha(g["y"][:400], draws, g["factors"])
    want = mle.numpy()
    assert ret[0][2] == (0, 4) and ret[1][2] == (4, 7)
    for r in range(world):
        pooled, full, _ = ret[r]
        assert abs(pooled - want.sum()) < 1e-9 * abs(want.sum())
        np.testing.assert_allclose(full, want, rtol=0, atol=0)
