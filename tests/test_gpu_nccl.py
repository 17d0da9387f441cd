"""The N>1 plumbing on one GPU: NCCL process group of world size 1 around pooled_objective."""
import os
import subprocess
import sys

import pytest
from conftest import ROOT

pytestmark = pytest.mark.gpu

SCRIPT = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
from metran_b200.batched import KalmanEngine
from metran_b200.distributed import pooled_objective, gather_per_model, shard_bounds
eng = KalmanEngine()
rng = np.random.default_rng(0)
B, T, m, k = 100, 60, 5, 3
y = rng.standard_normal((B, T, m)); alpha = rng.uniform(5, 50, (B, m + k)); G = rng.uniform(0.3, 0.8, (B, m, k)) / np.sqrt(k)
lo, hi = shard_bounds(B, 0, 1)
mle, _, _ = eng.dfm_loglik_alpha(y[lo:hi], alpha[lo:hi], G[lo:hi])
pooled = pooled_objective(eng, mle)
full = gather_per_model(mle, B)
torch.cuda.synchronize()
assert abs(float(pooled[0]) - float(mle.sum())) <= 1e-12 * abs(float(mle.sum()))
assert torch.equal(full, mle)
dist.destroy_process_group()
print("NCCL_OK")
"""


def test_pooled_objective_over_nccl_world_size_one():
    out = subprocess.run([sys.executable, "-c", SCRIPT % ROOT], capture_output=True, text=True, timeout=300)
    assert "NCCL_OK" in out.stdout, out.stderr[-3000:]
