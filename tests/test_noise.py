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


def check_noise_streams_and_cache(lib, device):
    """(1) Every search of a slot draws its root noise from its own Philox stream (the reference draws fresh noise at every root
    visit of every move, player.py:303-304): the same position searched twice must not see the same draws.  (2) The draws that
    K warps per game generate ahead of a wave (k_noise_fill) are a pure cache of the in-line sampler: switching the kernel off
    (CZ_NOISE_AHEAD=0) changes nothing — visit counts, W and the number of draws consumed agree bit for bit."""
    import os
    from oracle import senv as osenv
    from tests.search_checks import eval_planes
    res = {}
    for ahead in ("1", "0"):
        os.environ["CZ_NOISE_AHEAD"] = ahead
        try:
            eng = Engine(lib, device, n_games=3, sims_per_move=120, leaves_per_round=8, noise_mode=1, noise_eps=0.25, dirichlet_alpha=0.2,
                         c_puct=1.5, seed=11)
        finally:
            os.environ.pop("CZ_NOISE_AHEAD", None)
        eng.reset([osenv.INIT_STATE] * 3)
        out = []
        for rep in range(2):                       # the same root twice (no move played in between)
            eng.search_external(eval_planes, None)
            out.append([eng.root(g) for g in range(3)])
        res[ahead] = out
        eng.close()
    for rep in range(2):
        for g in range(3):
            a, b = res["1"][rep][g], res["0"][rep][g]
            assert a["n"] == b["n"] and a["w"] == b["w"] and a["noise_used"] == b["noise_used"] and a["noise_used"] > 0
    # the sampler itself: the stream of a slot changes with every search opened on it, and differs between slots
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
    eng.close()
    first, second = res["1"][0], res["1"][1]
    # different games of one search and the two searches of one game: different streams -> different visit distributions
    assert first[0]["n"] != first[1]["n"] or first[1]["n"] != first[2]["n"]
    assert any(first[g]["n"] != [x - y for x, y in zip(second[g]["n"], first[g]["n"])] for g in range(3))


def test_emul_noise_streams_and_cache(emul_lib):
    check_noise_streams_and_cache(emul_lib, "cpu")


@pytest.mark.gpu
def test_cuda_noise_streams_and_cache(cuda_lib):
    check_noise_streams_and_cache(cuda_lib, "cuda")
