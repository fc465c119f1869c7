"""Whole games of the REAL reference loops (tests/golden/games_k1.json.gz, recorded by oracle/gen_golden_games.py from the
unmodified SelfPlayWorker.start_game / EvaluateWorker.start_game at search_threads = 1):
  * the restated loops (oracle/selfplay.py, oracle/arena.py) replay every game move for move when they take their random
    decisions from the same generators -> pins the game-loop restatements (draw rules, repetition bans, resignation,
    final_move, value signs, store lottery) to the reference itself;
  * the ON-DEVICE game loop (cz_play_move) replays the games that contain no random decision."""
import gzip
import json
import os
import random

import numpy as np
import pytest

from oracle import arena as oarena
from oracle import player as op
from oracle import selfplay as osp
from tests.search_checks import eval_planes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _games():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "games_k1.json.gz"), "rt") as f:
        return json.load(f)["games"]


class _HostDraws:
    """The reference's own generators (what oracle/ref_worker_harness.ReferenceDraws does; restated here because the
    harness imports the reference tree, which is absent on the GPU box)."""

    def resign_lottery(self):
        return random.random()

    def store_lottery(self):
        return random.random()

    def playouts(self, lo, hi):
        return random.randint(lo, hi) * 100                 # evaluator.py:12,153 `from random import randint`

    def choose_with_player(self, player, state, turns, no_act, increase_temp):
        player.increase_temp = increase_temp
        policy, _ = player.calc_policy(state, turns, no_act)
        if no_act is not None:
            for act in no_act:
                policy[player.move_lookup[act]] = 0
        return player.labels[int(np.random.choice(range(len(player.labels)), p=player.apply_temperature(policy, turns)))]


def _pc(g):
    p = g["play"]
    return op.PlayConfig(simulation_num_per_move=g["sims"], search_threads=1, c_puct=p["c_puct"], noise_eps=p["noise_eps"],
                         dirichlet_alpha=p["dirichlet_alpha"], tau_decay_rate=p["tau_decay_rate"], virtual_loss=p["virtual_loss"],
                         resign_threshold=p["resign_threshold"], min_resign_turn=p["min_resign_turn"])


def test_restated_game_loops_replay_the_real_games():
    kinds = set()
    for g in _games():
        random.seed(g["seed"])
        np.random.seed(g["seed"])
        want = g["result"]
        d = _HostDraws()
        if g["kind"] == "selfplay":
            hist = bool(g.get("use_history"))
            r = osp.play_game(_pc(g), op.fake_evaluate_states_hist if hist else op.fake_evaluate_states, d,
                              max_game_length=g["play"]["max_game_length"], enable_resign_rate=g["play"]["enable_resign_rate"],
                              use_history=hist)
            assert (r["turns"], r["value_red"], r["store"], r["final_state"]) == \
                   (want["turns"], want["value_red"], want["store"], want["final_state"]), (g["seed"], g["sims"])
            if want["moves"] is not None:
                assert r["moves"] == want["moves"]
            kinds.add(("resign" if r["flags"] & 1 else "draw" if r["flags"] & 2 else "capture", want["store"]))
        else:
            real = not g.get("playouts_patched", True)       # the game ran with the reference's own `randint(8, 12) * 100`
            pc = _pc(g)
            if real:
                pc.simulation_num_per_move = -1              # must come from the draw
            r = oarena.play_arena_game(pc, op.fake_evaluate_states, op.fake_evaluate_states, g["idx"], lambda slot: d, 1,
                                       max_game_length=g["play"]["max_game_length"], playouts=(8, 12) if real else None)
            if real:
                assert r["playouts"] == want["playouts"] == g["sims"] and g["sims"] in (800, 900, 1000, 1100, 1200)
                kinds.add("real_playouts")
            assert (r["turns"], r["value_red"]) == (want["turns"], want["value_red"])
            assert r["moves"][:len(want["moves"])] == want["moves"] and len(r["moves"]) - len(want["moves"]) in (0, 1)
    assert {("resign", False), ("resign", True), ("draw", True), ("capture", True), "real_playouts"} <= kinds


def check_device_loop_replays_real_games(lib, device):
    from cczero_b200.engine import Engine

    def play(g, arena, want_records, slots=None, game_sims=None):
        p = g["play"]
        eng = Engine(lib, device, n_games=slots or (2 if arena else 1), sims_per_move=g["sims"] if game_sims is None else 7,
                     leaves_per_round=1, noise_mode=1,
                     noise_eps=0.0, c_puct=p["c_puct"], tau_decay_rate=0.0, max_game_length=p["max_game_length"],
                     resign_threshold=p["resign_threshold"], enable_resign_rate=0.0, min_resign_turn=p["min_resign_turn"], seed=1,
                     max_nodes_per_game=g["sims"] * 2 * p["max_game_length"] + 64, arena=arena,
                     use_history=bool(g.get("use_history")))
        eng.reset()
        if game_sims is not None:                            # per-game simulation_num_per_move (evaluator.py:153-154)
            eng.set_game_sims([game_sims] * eng.n_games)
        recs = []
        for _ in range(want_records * (2 * p["max_game_length"] + 4)):
            eng.search_external(eval_planes, None)
            if eng.play_move():
                recs += eng.drain_records()
            if len(recs) >= want_records:
                break
        assert int(eng.counters()[6]) == 0 and int(eng.counters()[4]) == 0
        eng.close()
        return recs
    det = [g for g in _games() if g["deterministic"]]
    sp = [g for g in det if g["kind"] == "selfplay"]
    assert len(sp) >= 3 and any(g.get("use_history") for g in sp)
    for g in sp:
        rec = play(g, False, 1)[0]
        want = g["result"]
        assert rec["moves"] == want["moves"], (g["seed"], rec["moves"][:6], want["moves"][:6])
        assert rec["value_red"] == want["value_red"] and rec["n_plies"] == want["turns"]
    ar = sorted((g for g in det if g["kind"] == "arena" and g.get("playouts_patched", True)), key=lambda g: g["idx"])
    assert [g["idx"] for g in ar[:2]] == [0, 1] and ar[0]["sims"] == ar[1]["sims"]
    recs = sorted(play(ar[0], True, 2), key=lambda r: r["game_index"])
    for rec, g in zip(recs, ar):
        want = g["result"]
        assert rec["game_index"] == g["idx"] and rec["value_red"] == want["value_red"] and rec["n_plies"] == want["turns"]
        assert rec["moves"][:len(want["moves"])] == want["moves"]
    # arena games that ran with the reference's own per-game `randint(8, 12) * 100` playouts: the engine default (7) is wrong on
    # purpose, the per-game value set through cz_set_game_sims must be what every search of the game runs
    real = [g for g in det if g["kind"] == "arena" and not g.get("playouts_patched", True)]
    assert len(real) >= 2 and {g["idx"] for g in real} == {0, 1}
    for g in real:
        m = g["idx"] + 1                                     # idx 1 is the second of two concurrent games
        recs = [r for r in play(g, True, m, slots=2 * m, game_sims=g["sims"]) if r["game_index"] == g["idx"]]
        want = g["result"]
        assert len(recs) == 1 and recs[0]["value_red"] == want["value_red"] and recs[0]["n_plies"] == want["turns"]
        assert recs[0]["moves"][:len(want["moves"])] == want["moves"], (g["seed"], recs[0]["moves"], want["moves"])


def test_emul_device_loop_replays_real_games(emul_lib):
    check_device_loop_replays_real_games(emul_lib, "cpu")


@pytest.mark.gpu
def test_cuda_device_loop_replays_real_games(cuda_lib):
    check_device_loop_replays_real_games(cuda_lib, "cuda")
