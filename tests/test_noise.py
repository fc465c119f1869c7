"""On-device root noise (noise_mode 1) is statistical parity only: its draws must be distributed like
np.random.dirichlet(alpha * ones(L))[0], i.e. Beta(alpha, (L-1) alpha)."""
import ctypes as C

import numpy as np
import pytest
import torch

from cczero_b200.engine import Engine


def check_noise(lib, device):
    eng = Engine(lib, device, n_games=2, sims_per_move=8, leaves_per_round=2, noise_mode=1, dirichlet_alpha=0.2, seed=7)
    n = 40000
    for L in (20, 44):
        out = torch.zeros(n, dtype=torch.float64, device=eng.device)
        lib.call("cz_noise_sample", eng._h, 1, L, n, C.c_void_p(out.data_ptr()))
        if eng.device.type == "cuda":
            torch.cuda.synchronize()
        x = out.cpu().numpy()
        a, b = 0.2, 0.2 * (L - 1)
        mean, var = a / (a + b), a * b / ((a + b) ** 2 * (a + b + 1))
        assert (x >= 0).all() and (x <= 1).all()
        assert abs(x.mean() - mean) < 5 * np.sqrt(var / n), (L, x.mean(), mean)
        assert abs(x.var() - var) < 0.1 * var, (L, x.var(), var)
        # compare a few quantiles with numpy's own sampler
        ref = np.random.RandomState(0).dirichlet(0.2 * np.ones(L), size=n)[:, 0]
        for q in (0.5, 0.8, 0.95):
            assert abs(np.quantile(x, q) - np.quantile(ref, q)) < 0.02 + 0.15 * np.quantile(ref, q), (L, q)
    eng.close()


def test_emul_noise_distribution(emul_lib):
    check_noise(emul_lib, "cpu")


@pytest.mark.gpu
def test_cuda_noise_distribution(cuda_lib):
    check_noise(cuda_lib, "cuda")


def check_noise_streams(lib, device):
    """Every search of a slot draws its root noise from its own Philox stream (the reference draws fresh noise at every root visit
    of every move, player.py:303-304); slots differ; the same search draws the same values again (a counter-based generator)."""
    from oracle import senv as osenv
    from tests.search_checks import eval_planes
    eng = Engine(lib, device, n_games=2, sims_per_move=8, leaves_per_round=2, noise_mode=1, dirichlet_alpha=0.2, seed=7)
    eng.reset([osenv.INIT_STATE] * 2)

    def draws(game):
        out = torch.zeros(64, dtype=torch.float64, device=eng.device)
        lib.call("cz_noise_sample", eng._h, game, 44, 64, C.c_void_p(out.data_ptr()))
        if eng.device.type == "cuda":
            torch.cuda.synchronize()
        return out.cpu().numpy().copy()
    d0 = draws(1)
    assert (d0 == draws(1)).all() and not (d0 == draws(0)).all()
    eng.search_external(eval_planes, None)
    d1 = draws(1)
    eng.search_external(eval_planes, None)
    d2 = draws(1)
    assert not (d0 == d1).any() and not (d1 == d2).any()
    assert eng.root(1)["noise_used"] > 0
    eng.close()


def test_emul_noise_streams(emul_lib):
    check_noise_streams(emul_lib, "cpu")


@pytest.mark.gpu
def test_cuda_noise_streams(cuda_lib):
    check_noise_streams(cuda_lib, "cuda")
