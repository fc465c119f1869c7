"""Root visit distributions of the REAL, threaded reference player (agent/player.py, search_threads = 10: a racy thread pool,
not reproducible run to run) -> tests/golden/mcts_k10_threaded.json.gz.  Build container only.

The GPU tier cannot import the reference, so the statistical comparison "canonical schedule at K = 10 vs the real threaded
player" (tests/test_oracle_vs_reference.py does it live against the oracle) travels as a fixture: per position, the visit
counts of `N_SEEDS` independent real searches.  tests/test_search.py compares the device engine's distributions with them.

    python -m oracle.gen_golden_k10
"""
import gzip
import json
import os

from . import senv as o
from .ref_player_harness import real_player_moves

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SIMS, K, N_SEEDS = 300, 10, 16


def positions():
    import random
    rng = random.Random(5)
    out = [o.INIT_STATE]
    for plies in (9, 24):
        s = o.INIT_STATE
        for _ in range(plies):
            s = o.step(s, rng.choice(o.get_legal_moves(s)))
        out.append(s)
    return out


def main():
    rows = []
    for s in positions():
        lm = o.get_legal_moves(s)
        runs = []
        for seed in range(N_SEEDS):
            r = real_player_moves([(s, 0, None, False)], SIMS, 1000 + seed, search_threads=K)[0]
            runs.append([int(r[1].get(m, (0,))[0]) for m in lm])
        rows.append({"state": s, "moves": lm, "visits": runs})
        print(s, [sum(v) for v in runs][:4])
    with gzip.open(os.path.join(GOLD, "mcts_k10_threaded.json.gz"), "wt") as f:
        json.dump({"generator": "oracle/gen_golden_k10.py", "reference": "agent/player.py CChessPlayer, search_threads=10, fake network "
                   "(oracle.player.fake_eval_from_planes) over a real Pipe", "sims": SIMS, "search_threads": K, "c_puct": 1.5,
                   "noise_eps": 0.25, "dirichlet_alpha": 0.2, "rows": rows}, f)


if __name__ == "__main__":
    main()
