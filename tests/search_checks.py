"""Shared search-parity checks (run against the CUDA library with -m gpu and against the CPU SIMT-emulation
build of the same kernel source otherwise)."""
import gzip
import json
import os
import threading
from multiprocessing import Pipe
from types import SimpleNamespace

import numpy as np

from cczero_b200.engine import Engine
from cczero_b200.player import CChessPlayer
from oracle import player as op
from oracle import senv as osenv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_mcts_golden(name="mcts_k1.json.gz"):
    with gzip.open(os.path.join(ROOT, "tests", "golden", name), "rt") as f:
        return json.load(f)


class FakeNetServer:
    """Stands in for CChessModelAPI.predict_batch_worker (api.py:37-74) with the deterministic pseudo-network."""

    def __init__(self):
        self.me, self.you = Pipe()
        self.stop = False
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        while not self.stop:
            if self.me.poll(0.001):
                try:
                    planes = self.me.recv()
                except EOFError:
                    return
                self.me.send([op.fake_eval_from_planes(p) for p in planes])

    def close(self):
        self.stop = True


def make_config(sims, k, **over):
    play = SimpleNamespace(simulation_num_per_move=sims, search_threads=k, c_puct=1.5, noise_eps=0.25, dirichlet_alpha=0.2,
                           tau_decay_rate=0.98, virtual_loss=3, resign_threshold=-0.92, min_resign_turn=20, max_game_length=100)
    for key, v in over.items():
        setattr(play, key, v)
    return SimpleNamespace(play=play, opts=SimpleNamespace(evaluate=False), model=None)


def eval_planes(planes):
    out = [op.fake_eval_from_planes(p) for p in planes]
    return np.stack([o[0] for o in out]), np.array([o[1] for o in out], dtype=np.float32)


def check_golden_k1(lib, device, use_history=False, name=None):
    """The drop-in CChessPlayer must reproduce the REAL reference player (search_threads=1): chosen move,
    N, W, P of every root edge, for every call of every golden case (tests/golden/mcts_k1.json.gz; with use_history the
    28-plane cases of mcts_k1_hist.json.gz, with and without the `hist` argument of action())."""
    gold = load_mcts_golden(name or ("mcts_k1_hist.json.gz" if use_history else "mcts_k1.json.gz"))
    eps = gold["config"].get("noise_eps", 0.25)
    for case in gold["cases"]:
        srv = FakeNetServer()
        np.random.seed(case["seed"])
        player = CChessPlayer(make_config(case["sims"], 1, noise_eps=eps), pipes=srv.you, lib=lib, device=device,
                              use_history=use_history)
        try:
            for call in case["calls"]:
                action, policy = player.action(call["state"], call["turns"], call["no_act"], increase_temp=call["increase_temp"],
                                               hist=call.get("hist"))
                root = player.engine.root(0)
                assert root["moves"] == call["legal"], case["name"]
                assert root["sum_n"] == call["sum_n"], (case["name"], root["sum_n"], call["sum_n"])
                for m, n, w, p in zip(root["moves"], root["n"], root["w"], root["p"]):
                    gn, gw, gq, gp = call["edges"].get(m, [0, 0.0, 0.0, 0.0])
                    assert n == gn and w == gw and float(np.float32(p)) == gp, (case["name"], m, (n, w, p), (gn, gw, gp))
                assert action == call["action"], (case["name"], action, call["action"])
                assert abs(sum(policy) - 1.0) < 1e-9
            # answering from the finished tree (close_and_return_action, player.py:88-106) is a legal move of the last state
            if "rand_after" in case:                     # the np.random stream stands where the reference left it
                assert float(np.random.rand()) == case["rand_after"]
            last = case["calls"][-1]
            got = player.close_and_return_action(last["state"], last["turns"], last["no_act"])
            assert got is not None and got[0] in last["legal"]
        finally:
            player.close()
            srv.close()


def oracle_search(state, sims, k, seed, no_act=None, increase_temp=False, eps=0.25, use_history=False, hist=None):
    pc = op.PlayConfig(simulation_num_per_move=sims, search_threads=k, c_puct=1.5, noise_eps=eps, dirichlet_alpha=0.2,
                       tau_decay_rate=0.98, virtual_loss=3)
    np.random.seed(seed)
    pl = op.OraclePlayer(pc, op.fake_evaluate_states_hist if use_history else op.fake_evaluate_states, use_history=use_history)
    pl.search(state, no_act, increase_temp, hist=hist)
    return pl


def engine_search(lib, device, states, sims, k, seed, no_act=None, increase_temp=False, eps=0.25, use_history=False, hists=None):
    """All games share the seed: game g searches states[g] with its own copy of the oracle's noise stream."""
    g = len(states)
    tables = []
    for s in states:
        L = len(osenv.get_legal_moves(s))
        np.random.seed(seed)
        tables.append([np.random.dirichlet(0.2 * np.ones(L))[0] for _ in range((sims + 40) * L)])
    width = max(len(t) for t in tables)
    noise = np.zeros((g, width))
    for i, t in enumerate(tables):
        noise[i, :len(t)] = t
    eng = Engine(lib, device, n_games=g, sims_per_move=sims, leaves_per_round=k, noise_mode=0, c_puct=1.5, noise_eps=eps,
                 dirichlet_alpha=0.2, tau_decay_rate=0.98, use_history=use_history)
    eng.reset(states)
    stats = eng.search_external(eval_planes, eng.make_opts(no_act=[no_act] * g if no_act else None,
                                                           increase_temp=[1 if increase_temp else 0] * g, noise=noise,
                                                           hist=hists))
    return eng, stats


def compare_root(eng, game, pl, state):
    node = pl.tree[state]
    r = eng.root(game)
    assert r["moves"] == node.legal_moves
    assert r["sum_n"] == node.sum_n
    for m, n, w, p in zip(r["moves"], r["n"], r["w"], r["p"]):
        e = node.a.get(m)
        en, ew, ep = (e.n, float(e.w), float(e.p)) if e is not None else (0, 0.0, 0.0)
        assert (n, w, float(np.float32(p))) == (en, ew, ep), (state, m, (n, w, p), (en, ew, ep))
    assert r["noise_used"] == pl.stats["noise_draws"]
    assert r["sims_run"] == pl.stats["sims"]
    assert int(eng.counters()[6]) == 0 and int(eng.counters()[4]) == 0      # no error flag, no table reset


def midgame_states(n, seed, lo=15, hi=80):
    rng = np.random.RandomState(seed)
    out = []
    while len(out) < n:
        s = osenv.INIT_STATE
        ok = True
        for _ in range(rng.randint(lo, hi)):
            if osenv.done(s)[0]:
                ok = False
                break
            lm = osenv.get_legal_moves(s)
            s = osenv.step(s, lm[rng.randint(len(lm))])
        if ok and not osenv.done(s)[0]:
            out.append(s)
    return out


def game_history(plies, seed):
    """[s0, a0, s1, ..., s_plies] of a random playout that is not over (the `history` list of worker/self_play.py:111-147)."""
    rng = np.random.RandomState(seed)
    while True:
        s = osenv.INIT_STATE
        hist = [s]
        for _ in range(plies):
            if osenv.done(s)[0]:
                break
            lm = osenv.get_legal_moves(s)
            a = lm[rng.randint(len(lm))]
            s = osenv.step(s, a)
            hist += [a, s]
        if len(hist) == 2 * plies + 1 and not osenv.done(s)[0]:
            return hist


def check_history_vs_oracle(lib, device, cases=((90, 1, 1), (160, 8, 2), (200, 16, 3))):
    """use_history (28 planes): engine == oracle at any K, games with a long `hist`, a short one (< 5 entries) and none."""
    for (sims, k, seed) in cases:
        hists = [game_history(24 + seed, 50 + seed), game_history(1, 60 + seed), None, game_history(37, 70 + seed)]
        states = [hists[0][-1], hists[1][-1], midgame_states(1, 80 + seed)[0], hists[3][-1]]
        eng, _ = engine_search(lib, device, states, sims, k, seed, use_history=True, hists=hists)
        for g, s in enumerate(states):
            compare_root(eng, g, oracle_search(s, sims, k, seed, use_history=True, hist=hists[g]), s)
        eng.close()


def check_search_fuzz(lib, device, n_cases=10, seed=2024):
    """Seeded random configurations (simulations, K, c_puct, noise_eps, virtual loss) on random positions - reachable
    mid-games, sparse endgames and arbitrary boards with up to ~80 legal moves (three 32-lane chunks in select): engine ==
    oracle bit for bit."""
    from tests.env_checks import random_boards
    rng = np.random.RandomState(seed)
    pool = midgame_states(6, seed % 1000, lo=5, hi=150) + [s for s in random_boards(60, seed, own="RRKKCCPPPPP") if not osenv.done(s)[0]][:8] + \
        ['3s5/9/9/9/9/9/9/9/9/4S4', '5s3/9/9/9/9/9/9/9/5r3/3S5'] + \
        ['R2msm2R/4m4/9/C7C/9/9/C7C/9/9/R3S3R', 'R2msm2R/4m4/9/C7C/9/4R4/C7C/9/9/R3S3R',        # 67 and 82 legal moves
         'R2msm2R/4m4/9/C7C/1R5R1/9/C7C/9/9/R3S3R']                                                # 97
    assert max(len(osenv.get_legal_moves(s)) for s in pool) > 96
    for case in range(n_cases):
        sims, k = int(rng.choice([17, 40, 75, 130])), int(rng.choice([1, 2, 3, 8, 16, 40]))
        c_puct, eps, vl = float(rng.choice([0.5, 1.5, 5.0])), float(rng.choice([0.0, 0.25, 1.0])), int(rng.choice([1, 3]))
        states = [pool[i] for i in rng.choice(len(pool), size=2, replace=False)] + [pool[-1 - case % 3]]
        tables = []
        for s in states:
            L = len(osenv.get_legal_moves(s))
            np.random.seed(case)
            tables.append([np.random.dirichlet(0.2 * np.ones(L))[0] for _ in range((sims + 2 * k + 2) * L)])
        noise = np.zeros((len(states), max(len(t) for t in tables)))
        for i, t in enumerate(tables):
            noise[i, :len(t)] = t
        eng = Engine(lib, device, n_games=len(states), sims_per_move=sims, leaves_per_round=k, virtual_loss=vl, noise_mode=0,
                     c_puct=c_puct, noise_eps=eps, dirichlet_alpha=0.2, tau_decay_rate=0.98)
        eng.reset(states)
        eng.search_external(eval_planes, eng.make_opts(noise=noise))
        for g, s in enumerate(states):
            pc = op.PlayConfig(simulation_num_per_move=sims, search_threads=k, c_puct=c_puct, noise_eps=eps, dirichlet_alpha=0.2,
                               tau_decay_rate=0.98, virtual_loss=vl)
            np.random.seed(case)
            pl = op.OraclePlayer(pc, op.fake_evaluate_states)
            pl.search(s)
            compare_root(eng, g, pl, s)
        eng.close()


def check_vs_oracle(lib, device, cases):
    """Canonical K-round schedule: engine == oracle restatement bit for bit (N, W, P, sum_n, noise draws)."""
    for (sims, k, seed, n_states) in cases:
        states = [osenv.INIT_STATE] + midgame_states(n_states - 1, seed)
        eng, stats = engine_search(lib, device, states, sims, k, seed)
        tot = {"sims": 0, "path_edges": 0, "no_network": 0, "positions": 0}
        for g, s in enumerate(states):
            pl = oracle_search(s, sims, k, seed)
            compare_root(eng, g, pl, s)
            for key in tot:
                tot[key] += pl.stats[key]
            # the most visited line (print_depth_info, player.py:408-450) read back through cz_get_pv
            _, _, pv = pl.depth_info(s, 0, 0.0, None)
            moves, _ = eng.pv(g, 20)
            assert "".join(" " + osenv.to_uci_move(osenv.flip_move(m) if i % 2 else m) for i, m in enumerate(moves)) == pv
        st = eng.search_stats()                          # the counters behind bench.py's byte-model figures
        assert (st["sims"], st["path_edges"], st["no_network"], st["nodes_created"]) == \
               (tot["sims"], tot["path_edges"], tot["no_network"], tot["positions"]), (st, tot)
        eng.close()


def check_no_act_and_temp(lib, device):
    s = midgame_states(1, 99)[0]
    ban = osenv.get_legal_moves(s)[:3]
    eng, _ = engine_search(lib, device, [s], 120, 4, 5, no_act=ban)
    compare_root(eng, 0, oracle_search(s, 120, 4, 5, no_act=ban), s)
    eng.close()
    eng, _ = engine_search(lib, device, [s], 90, 8, 6, increase_temp=True)
    compare_root(eng, 0, oracle_search(s, 90, 8, 6, increase_temp=True), s)
    eng.close()


def check_terminal_and_repetition(lib, device):
    """Positions one or two plies from a king capture, and long searches that revisit positions in-path."""
    # a bare-kings-plus-rooks ending: lots of checks, captures and repetitions inside the search
    endings = ['3s5/9/9/9/4r4/9/9/4R4/9/4S4', '4s4/4m4/9/9/9/9/2R6/9/4M4/3S1r3',
               '2e1s4/4m4/4e4/9/9/9/9/4C4/4M4/3S5']
    for i, s in enumerate(endings):
        if osenv.done(s)[0]:
            continue
        for (sims, k) in ((300, 1), (400, 8)):
            eng, _ = engine_search(lib, device, [s], sims, k, 40 + i, eps=0.25)
            compare_root(eng, 0, oracle_search(s, sims, k, 40 + i), s)
            eng.close()


def check_multi_move_reuse_and_options(lib, device):
    """Several plies with ONE engine / ONE oracle player (tree reuse, player.py:153-158) at K = 8, the `depth` argument
    of action() (cz_root_opts.sims_override, player.py:160-161) and the `active` mask."""
    sims, k = 120, 8
    pc = op.PlayConfig(simulation_num_per_move=sims, search_threads=k, c_puct=1.5, noise_eps=0.0, dirichlet_alpha=0.2,
                       tau_decay_rate=0.98, virtual_loss=3)
    pl = op.OraclePlayer(pc, op.fake_evaluate_states, noise=lambda n: 0.0)
    eng = Engine(lib, device, n_games=2, sims_per_move=sims, leaves_per_round=k, noise_mode=0, c_puct=1.5, noise_eps=0.0,
                 tau_decay_rate=0.98)
    other = midgame_states(1, 17)[0]
    eng.reset([osenv.INIT_STATE, other])
    state = osenv.INIT_STATE
    for ply in range(4):
        depth = 60 if ply == 2 else None                      # one ply searched with action(depth=60)
        pl.stats["sims"] = 0
        pl.stats["noise_draws"] = 0
        pl.search(state, depth=depth)
        eng.set_root(0, state)
        eng.search_external(eval_planes, eng.make_opts(active=[1, 0], sims_override=depth or 0))
        compare_root(eng, 0, pl, state)
        assert eng.root(1)["moves"] == [] and eng.root(1)["sum_n"] == 0    # the inactive game was never touched
        node = pl.tree[state]
        best = max(node.legal_moves, key=lambda m: (node.a[m].n if m in node.a else 0))
        state = osenv.step(state, best)
    # now the second game alone
    eng.search_external(eval_planes, eng.make_opts(active=[0, 1]))
    pl2 = op.OraclePlayer(pc, op.fake_evaluate_states, noise=lambda n: 0.0)
    pl2.search(other)
    compare_root(eng, 1, pl2, other)
    eng.close()
