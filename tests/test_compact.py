"""Pool compaction (cz_compact / automatic in cz_search_begin): kept nodes keep their statistics, the search goes on."""
import numpy as np
import pytest

from cczero_b200.engine import Engine
from oracle import senv as osenv
from tests.search_checks import eval_planes


def _engine(lib, device, nodes, sims=24, k=4, games=3):
    return Engine(lib, device, n_games=games, sims_per_move=sims, leaves_per_round=k, noise_mode=1, noise_eps=0.0,
                  tau_decay_rate=0.9, max_game_length=40, seed=21, max_nodes_per_game=nodes, max_edges_per_game=nodes * 48)


def check_compact(lib, device):
    a, b = _engine(lib, device, 4096), _engine(lib, device, 4096)
    a.reset()
    b.reset()
    for _ in range(4):
        for e in (a, b):
            e.search_external(eval_planes, None)
            e.play_move()
    b.compact()
    assert int(b.counters()[5]) == 3 and int(b.counters()[4]) == 0
    for g in range(3):
        ra, rb = a.root(g), b.root(g)
        assert ra == rb and ra["sum_n"] >= 1                 # the subtree under the new root survived, bit for bit
    # the search continues from the compacted table and reuses it (fewer than `sims` simulations needed)
    b.search_external(eval_planes, None)
    assert (b.sims_run() < 24).all() and (b.sims_run() > 0).all()
    for g in range(3):
        r = b.root(g)
        assert r["sum_n"] == 24 and sum(r["n"]) == 23
    a.close()
    b.close()
    # automatic: pools far too small for a whole game -> compaction instead of dropping the table
    c = _engine(lib, device, 64, sims=24, k=4, games=2)
    c.reset()
    recs = []
    for _ in range(30):
        c.search_external(eval_planes, None)
        if c.play_move():
            recs += c.drain_records()
    cnt = c.counters()
    assert int(cnt[5]) > 0
    for g in range(2):
        assert c.root(g)["sum_n"] >= 0
    for r in recs:
        s = osenv.INIT_STATE
        for m in r["moves"]:
            assert m in osenv.get_legal_moves(s)
            s = osenv.step(s, m)
    c.close()


def test_emul_compact(emul_lib):
    check_compact(emul_lib, "cpu")


@pytest.mark.gpu
def test_cuda_compact(cuda_lib):
    check_compact(cuda_lib, "cuda")


def check_overflow_paths(lib, device):
    """Pools / path storage too small for the search: the engine must flag it (counters[6]) and keep going, never corrupt."""
    # node pool of 16 for a 60-simulation search: nodes run out inside the search -> flag 2, the rest finish as draws
    e = Engine(lib, device, n_games=2, sims_per_move=60, leaves_per_round=4, noise_mode=1, noise_eps=0.0, seed=1,
               max_nodes_per_game=16, max_edges_per_game=16 * 48)
    e.reset()
    e.search_external(eval_planes, None)
    c = e.counters()
    assert int(c[6]) & 2 and int(c[7]) == 2
    for g in range(2):
        r = e.root(g)
        assert r["sims_run"] == 60 and r["sum_n"] >= 1 and sum(r["n"]) <= 60
    e.play_move()                                             # still able to pick a move and continue
    e.reset()
    assert int(e.counters()[6]) == 0                          # flags clear with the games
    e.close()
    # path storage of 8 plies with a deep, narrow search (one legal line is forced by banning everything else is not
    # possible here, so just search long enough for some line to exceed 8 plies)
    e = Engine(lib, device, n_games=1, sims_per_move=1500, leaves_per_round=8, noise_mode=1, noise_eps=0.0, seed=2, max_path=8,
               c_puct=0.05, max_nodes_per_game=8192)
    e.reset(['3s5/9/9/9/4r4/9/9/4R4/9/4S4'])
    e.search_external(eval_planes, None)
    r = e.root(0)
    assert r["sims_run"] == 1500
    assert int(e.counters()[6]) in (0, 1)                     # 1 = some simulation hit the path limit and was closed as a draw
    e.close()


def test_emul_overflow_paths(emul_lib):
    check_overflow_paths(emul_lib, "cpu")


@pytest.mark.gpu
def test_cuda_overflow_paths(cuda_lib):
    check_overflow_paths(cuda_lib, "cuda")
