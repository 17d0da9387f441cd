"""CPU-side checks of bench.py: helper arithmetic and the `--impl reference` arm's JSON line."""
import json
import os
import subprocess
import sys

import numpy as np
from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_flops_formula():
    # F_f(n, m) = 5 n^2 + m (5 n^2 + 7 n + 7)
    assert bench.flops_per_step(8, 5) == 2235
    assert bench.flops_per_step(6, 5) == 1325
    assert bench.flops_per_step(32, 64) == 347584


def test_synthetic_inputs_are_seeded_and_model_consistent():
    rng = np.random.default_rng(3)
    alpha, gamma = bench.make_params(rng, 7, 5, 3)
    assert alpha.shape == (7, 8) and gamma.shape == (7, 5, 3)
    assert np.all((alpha >= 5) & (alpha <= 50)) and np.all(np.sum(gamma ** 2, axis=2) < 1.0)   # communality < 1
    y1 = bench.simulate_cpu(np.random.default_rng(4), alpha, gamma, 50, 0.3)
    y2 = bench.simulate_cpu(np.random.default_rng(4), alpha, gamma, 50, 0.3)
    assert np.array_equal(np.isnan(y1), np.isnan(y2)) and np.allclose(np.nan_to_num(y1), np.nan_to_num(y2))
    assert 0.2 < np.isnan(y1).mean() < 0.4


def test_usable_cpus_reports_affinity():
    n, quota = bench.usable_cpus()
    assert n == len(os.sched_getaffinity(0)) and (quota is None or quota > 0)


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "3", "--config", "c1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == bench.UNIT and line["dtype"] == "f64"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 1e5
