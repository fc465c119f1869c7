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
