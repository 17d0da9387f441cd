"""The host-side mirrors (SPKalmanFilter, DynamicFactorModel, filtermethod contract) on the
REAL CUDA engine: same assertions as tests/test_host.py, which drives them through the
oracle-backed fake library on CPU."""
import numpy as np
import pytest

from test_host import (test_dfm_mask_unmask, test_dfm_objective_and_batch,  # noqa: F401
                       test_dfm_solve_reaches_reference_optimum, test_engine_error_behaviour,
                       test_filtermethod_contract, test_spkalmanfilter_mirror_on_b21)

pytestmark = pytest.mark.gpu


@pytest.fixture()
def eng():
    from metran_b200.batched import KalmanEngine

    return KalmanEngine()


def test_b21_solve_on_gpu_matches_notebook(eng, b21):
    """examples/example_script.py: -2logL at the optimum 2332.33 (notebook cell 8)."""
    from metran_b200.dfm import DynamicFactorModel

    mdl = DynamicFactorModel(b21["y"], b21["factors"], engine=eng)
    ok, x, stderr = mdl.solve()
    assert ok
    assert round(mdl.parameters["obj_func"], 2) == 2332.33
    assert abs(mdl.parameters["obj_func"] - float(b21["obj_opt"])) < 1e-5
    np.testing.assert_allclose(x, b21["alpha_opt"], rtol=2e-3)
    # unbatched objective path gives the same optimum
    mdl2 = DynamicFactorModel(b21["y"], b21["factors"], engine=eng)
    mdl2.solve(batched=False)
    assert abs(mdl2.parameters["obj_func"] - mdl.parameters["obj_func"]) < 1e-6


def test_host_streamed_api_matches_resident(eng):
    import torch

    rng = np.random.default_rng(5)
    B, T, m, k = 700, 96, 5, 3
    y = rng.standard_normal((B, T, m))
    y[rng.random(y.shape) < 0.1] = np.nan
    alpha = rng.uniform(5, 50, (B, m + k))
    G = rng.uniform(0.3, 0.9, (B, m, k)) / np.sqrt(k)
    want, wn, _ = eng.dfm_loglik_alpha(y, alpha, G)
    yh = torch.as_tensor(y).pin_memory()
    got, gn, gs = eng.dfm_loglik_alpha_host(yh, alpha, G, chunk_models=256)
    assert torch.equal(got, want.cpu()) and torch.equal(gn, wn.cpu())
    assert eng.probe_fp64_tflops(reps=1) > 1.0


def test_batched_fit_on_gpu(eng):
    """512 models fitted in lock step; spot-check three against the scipy single-model path."""
    import bench
    import torch

    from metran_b200.dfm import DynamicFactorModel
    from metran_b200.fit import fit_dfm_batch

    rng = np.random.default_rng(42)
    B, T, m, k = 512, 400, 5, 2
    alpha, gamma = bench.make_params(rng, B, m, k)
    y = bench.simulate_gpu(torch, eng.device, 7, eng.tensor(alpha), eng.tensor(gamma), T, 0.1)
    res = fit_dfm_batch(eng, y, gamma)
    assert res.converged.mean() > 0.97
    at_truth, _, _ = eng.dfm_loglik_alpha(y, alpha, gamma)
    assert np.all(res.obj <= at_truth.cpu().numpy() + 1e-6)          # the fit beats the generating parameters
    yh = y.cpu().numpy()
    for b in (0, 255, 511):
        mdl = DynamicFactorModel(yh[b], gamma[b], engine=eng)
        mdl.solve()
        assert res.obj[b] <= mdl.parameters["obj_func"] + 1e-3 * max(1.0, abs(mdl.parameters["obj_func"]))
