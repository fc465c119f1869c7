"""Whole games of the REAL reference loops (worker/self_play.py:95-212 SelfPlayWorker.start_game, worker/evaluator.py:
147-250 EvaluateWorker.start_game; search_threads = 1, deterministic fake network) -> tests/golden/games_k1.json.gz.
Build-container only (python -m oracle.gen_golden games).

Two kinds of use:
  * every game is replayed by the restated loops (oracle/selfplay.py, oracle/arena.py) taking their random decisions
    from the same generators (ref_worker_harness.ReferenceDraws): moves, result, length, store flag must be identical;
  * games flagged "deterministic" (tau_decay_rate = 0 -> arg-max moves, noise_eps = 0, no repetition that raises the
    temperature, >= 10 plies) do not depend on any random draw, so the ON-DEVICE game loop (cz_play_move) must replay
    them move for move as well.
"""
import gzip
import json
import os

from .ref_worker_harness import real_arena_game, real_selfplay_game

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SELFPLAY = [
    # (seed, sims, play-config overrides)
    (1, 30, dict(max_game_length=12)),
    (4, 20, dict(max_game_length=60)),
    (5, 20, dict(max_game_length=60, resign_threshold=-0.02, min_resign_turn=2, enable_resign_rate=0.0)),
    (9, 12, dict(max_game_length=100)),
    (10, 16, dict(max_game_length=60, resign_threshold=0.3, min_resign_turn=2, enable_resign_rate=0.0)),
    (13, 16, dict(max_game_length=60, resign_threshold=0.2, min_resign_turn=3, enable_resign_rate=0.0)),
    (7, 24, dict(max_game_length=60, tau_decay_rate=0.0, noise_eps=0.0, enable_resign_rate=0.0)),
    (8, 24, dict(max_game_length=100, tau_decay_rate=0.0, noise_eps=0.0, enable_resign_rate=0.0)),
    (14, 24, dict(max_game_length=100, tau_decay_rate=0.0, noise_eps=0.0, enable_resign_rate=0.0, resign_threshold=0.15,
                  min_resign_turn=20)),
]
ARENA = [
    # (seed, idx, sims, overrides)
    (1, 0, 20, dict(max_game_length=40)),
    (2, 1, 20, dict(max_game_length=40)),
    (3, 0, 24, dict(max_game_length=100, tau_decay_rate=0.0, noise_eps=0.0)),
    (4, 1, 24, dict(max_game_length=100, tau_decay_rate=0.0, noise_eps=0.0)),
    (5, 0, 16, dict(max_game_length=100, noise_eps=0.0)),
]
# arena games run EXACTLY as written: `playouts = randint(8, 12) * 100` (evaluator.py:153-154) from the seeded `random`
# module sets the simulations per move of the game.  Short by max_game_length (800-1200 pure-Python simulations per ply).
ARENA_REAL = [
    # (seed, idx, overrides)
    (11, 0, dict(max_game_length=5, tau_decay_rate=0.0, noise_eps=0.0)),
    (12, 1, dict(max_game_length=4, tau_decay_rate=0.0, noise_eps=0.0)),
    (17, 0, dict(max_game_length=3)),
]
BASE = dict(tau_decay_rate=0.98, noise_eps=0.25, enable_resign_rate=0.1, resign_threshold=-0.5, min_resign_turn=4, c_puct=1.5,
            dirichlet_alpha=0.2, virtual_loss=3)


def _deterministic(play, game):
    return play["tau_decay_rate"] == 0.0 and play["noise_eps"] == 0.0 and not game["increase_temp_used"] and game["turns"] >= 10


def gen_games():
    games = []
    for seed, sims, over in SELFPLAY:
        play = dict(BASE, **over)
        g = real_selfplay_game(seed, sims, **play)
        games.append({"kind": "selfplay", "seed": seed, "sims": sims, "play": play, "result": g, "deterministic": _deterministic(play, g)})
    # use_history = True: 28-plane leaves (the game loop passes no `hist`, self_play.py:124)
    for seed, sims, over in ((21, 20, dict(max_game_length=40)),
                             (22, 22, dict(max_game_length=100, tau_decay_rate=0.0, noise_eps=0.0, enable_resign_rate=0.0))):
        play = dict(BASE, **over)
        g = real_selfplay_game(seed, sims, use_history=True, **play)
        games.append({"kind": "selfplay", "seed": seed, "sims": sims, "play": play, "result": g, "use_history": True,
                      "deterministic": _deterministic(play, g)})
    for seed, idx, sims, over in ARENA:
        play = dict(BASE, **over)
        g = real_arena_game(seed, idx, sims, **play)
        games.append({"kind": "arena", "seed": seed, "idx": idx, "sims": sims, "play": play, "result": g,
                      "deterministic": _deterministic(play, g), "playouts_patched": True})
    for seed, idx, over in ARENA_REAL:
        play = dict(BASE, **over)
        g = real_arena_game(seed, idx, None, **play)
        assert g["playouts"] in (800, 900, 1000, 1100, 1200)
        games.append({"kind": "arena", "seed": seed, "idx": idx, "sims": g["playouts"], "play": play, "result": g,
                      "deterministic": play["tau_decay_rate"] == 0.0 and play["noise_eps"] == 0.0 and not g["increase_temp_used"],
                      "playouts_patched": False})
    # deterministic arena games (no repetition): scan seeds-independent settings (the game does not depend on the seed)
    for sims in (18, 20, 22, 26, 28, 30, 34):
        if sum(1 for x in games if x["kind"] == "arena" and x["deterministic"] and x["playouts_patched"]) >= 2:
            break
        play = dict(BASE, max_game_length=100, tau_decay_rate=0.0, noise_eps=0.0)
        for idx in (0, 1):
            g = real_arena_game(0, idx, sims, **play)
            if _deterministic(play, g):
                games.append({"kind": "arena", "seed": 0, "idx": idx, "sims": sims, "play": play, "result": g, "deterministic": True,
                              "playouts_patched": True})
    out = {"generator": "oracle/gen_golden_games.py", "reference": "NeymarL/ChineseChess-AlphaZero @7f45b0c worker/self_play.py, worker/evaluator.py",
           "search_threads": 1, "games": games}
    with gzip.open(os.path.join(GOLD, "games_k1.json.gz"), "wt") as f:
        json.dump(out, f)
    for g in games:
        r = g["result"]
        print(g["kind"], g["seed"], g.get("idx"), g["sims"], "turns", r["turns"], "v", r["value_red"], "store", r.get("store"),
              "inc", r["increase_temp_used"], "det", g["deterministic"])
