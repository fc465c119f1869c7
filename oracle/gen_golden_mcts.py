"""K=1 seeded searches of the REAL reference CChessPlayer (fake deterministic network) -> tests/golden/mcts_k1.json.gz.
Build-container only.  Each case is a sequence of action() calls on ONE player object (tree reuse across moves)."""
import gzip
import json
import os
import random

from . import ref_import
from .ref_player_harness import real_player_moves

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _midgame_states(n, seed):
    r = ref_import.senv()
    rng = random.Random(seed)
    out = []
    while len(out) < n:
        s = r.INIT_STATE
        plies = rng.randint(20, 70)
        ok = True
        for _ in range(plies):
            if r.done(s)[0]:
                ok = False
                break
            s = r.step(s, rng.choice(r.get_legal_moves(s)))
        if ok and not r.done(s)[0]:
            out.append(s)
    return out


def gen_mcts():
    r = ref_import.senv()
    cases = []
    init = r.INIT_STATE
    mids = _midgame_states(4, 7)
    specs = [
        dict(name="init_60", seed=0, sims=60, calls=[(init, 0, None, False)]),
        dict(name="init_250", seed=3, sims=250, calls=[(init, 0, None, False)]),
        dict(name="mid0_200", seed=5, sims=200, calls=[(mids[0], 31, None, False)]),
        dict(name="mid1_400", seed=11, sims=400, calls=[(mids[1], 44, None, False)]),
        dict(name="mid2_no_act", seed=13, sims=150, calls=[(mids[2], 40, "FIRST2", False)]),
        dict(name="mid3_inc_temp", seed=17, sims=150, calls=[(mids[3], 12, None, True)]),
    ]
    for sp in specs:
        calls = []
        for (s, t, na, inc) in sp["calls"]:
            if na == "FIRST2":
                na = r.get_legal_moves(s)[:2]
            calls.append((s, t, na, inc))
        res = real_player_moves(calls, sp["sims"], sp["seed"])
        cases.append({"name": sp["name"], "seed": sp["seed"], "sims": sp["sims"],
                      "calls": [{"state": c[0], "turns": c[1], "no_act": c[2], "increase_temp": c[3],
                                 "action": a, "sum_n": sn, "legal": r.get_legal_moves(c[0]),
                                 "edges": {m: list(v) for m, v in e.items()}}
                                for c, (a, e, sn) in zip(calls, res)]})
    # one player, three consecutive plies of a game (tree reuse, player.py:153-158)
    seed, sims = 23, 120
    import numpy as np
    calls, s, t = [], init, 0
    pm = ref_import.player_module()
    from .ref_player_harness import FakeNetServer, make_config
    cfg = make_config(sims, 1)
    srv = FakeNetServer()
    np.random.seed(seed)
    player = pm.CChessPlayer(cfg, pipes=srv.you, enable_resign=False)
    seq = []
    for ply in range(3):
        a, _ = player.action(s, t)
        node = player.tree[s]
        seq.append({"state": s, "turns": t, "no_act": None, "increase_temp": False, "action": a, "sum_n": int(node.sum_n),
                    "legal": r.get_legal_moves(s),
                    "edges": {m: [int(x.n), float(x.w), float(x.q), float(x.p)] for m, x in node.a.items()}})
        s = r.step(s, a)
        t += 1
    player.close(wait=False)
    srv.close()
    cases.append({"name": "three_plies_reuse", "seed": seed, "sims": sims, "calls": seq})
    out = {"generator": "oracle/gen_golden_mcts.py", "reference": "NeymarL/ChineseChess-AlphaZero @7f45b0c agent/player.py",
           "config": {"search_threads": 1, "c_puct": 1.5, "noise_eps": 0.25, "dirichlet_alpha": 0.2, "tau_decay_rate": 0.98,
                      "virtual_loss": 3}, "cases": cases}
    with gzip.open(os.path.join(GOLD, "mcts_k1.json.gz"), "wt") as f:
        json.dump(out, f)
    print("mcts cases:", [(c["name"], [x["action"] for x in c["calls"]]) for c in cases])


def _game_history(plies, seed):
    """[s0, a0, s1, ..., s_plies] of a random playout that is not over."""
    r = ref_import.senv()
    rng = random.Random(seed)
    while True:
        s = r.INIT_STATE
        hist = [s]
        for _ in range(plies):
            if r.done(s)[0]:
                break
            a = rng.choice(r.get_legal_moves(s))
            s = r.step(s, a)
            hist += [a, s]
        if len(hist) == 2 * plies + 1 and not r.done(s)[0]:
            return hist


def gen_mcts_hist():
    """use_history=True (28 input planes, static_env.py:158-194 / player.py:326-334): plane vectors and K=1 searches of the
    real player, with and without the `hist` argument of action() -> tests/golden/mcts_k1_hist.json.gz."""
    import numpy as np
    r = ref_import.senv()
    planes = []
    for seed, plies in ((1, 1), (2, 2), (3, 7), (4, 30), (5, 55)):
        h = _game_history(plies, seed)
        p = r.state_history_to_planes(h[-1], h)
        planes.append({"state": h[-1], "history": h[-5:], "nonzero": np.flatnonzero(p.reshape(-1)).tolist()})
    h30, h41, h3 = _game_history(30, 11), _game_history(41, 12), _game_history(1, 13)
    specs = [
        dict(name="hist_init_80_nohist", seed=2, sims=80, calls=[(r.INIT_STATE, 0, None, False, None)]),
        dict(name="hist_mid30_150", seed=4, sims=150, calls=[(h30[-1], 30, None, False, h30)]),
        dict(name="hist_mid41_nohist_200", seed=6, sims=200, calls=[(h41[-1], 41, None, False, None)]),
        dict(name="hist_short_60", seed=8, sims=60, calls=[(h3[-1], 1, None, False, h3)]),
        dict(name="hist_mid41_no_act", seed=9, sims=120, calls=[(h41[-1], 41, "FIRST2", False, h41)]),
    ]
    cases = []
    for sp in specs:
        calls = []
        for (s, t, na, inc, hi) in sp["calls"]:
            if na == "FIRST2":
                na = r.get_legal_moves(s)[:2]
            calls.append((s, t, na, inc, hi))
        res = real_player_moves(calls, sp["sims"], sp["seed"], use_history=True)
        cases.append({"name": sp["name"], "seed": sp["seed"], "sims": sp["sims"],
                      "calls": [{"state": c[0], "turns": c[1], "no_act": c[2], "increase_temp": c[3], "hist": c[4],
                                 "action": a, "sum_n": sn, "legal": r.get_legal_moves(c[0]),
                                 "edges": {m: list(v) for m, v in e.items()}}
                                for c, (a, e, sn) in zip(calls, res)]})
    # one player over three plies of a game with the growing game history passed in (uci.py:288)
    from .ref_player_harness import FakeNetServer, make_config
    pm = ref_import.player_module()
    seed, sims = 29, 100
    cfg = make_config(sims, 1)
    srv = FakeNetServer()
    np.random.seed(seed)
    player = pm.CChessPlayer(cfg, pipes=srv.you, enable_resign=False, use_history=True)
    hist = list(_game_history(6, 21))
    s, t, seq = hist[-1], 6, []
    for ply in range(3):
        a, _ = player.action(s, t, hist=list(hist))
        node = player.tree[s]
        seq.append({"state": s, "turns": t, "no_act": None, "increase_temp": False, "hist": list(hist), "action": a,
                    "sum_n": int(node.sum_n), "legal": r.get_legal_moves(s),
                    "edges": {m: [int(x.n), float(x.w), float(x.q), float(x.p)] for m, x in node.a.items()}})
        s = r.step(s, a)
        hist += [a, s]
        t += 1
    player.close(wait=False)
    srv.close()
    cases.append({"name": "hist_three_plies_reuse", "seed": seed, "sims": sims, "calls": seq})
    out = {"generator": "oracle/gen_golden_mcts.py:gen_mcts_hist", "reference": "NeymarL/ChineseChess-AlphaZero @7f45b0c agent/player.py",
           "config": {"search_threads": 1, "c_puct": 1.5, "noise_eps": 0.25, "dirichlet_alpha": 0.2, "tau_decay_rate": 0.98,
                      "virtual_loss": 3, "use_history": True}, "planes": planes, "cases": cases}
    with gzip.open(os.path.join(GOLD, "mcts_k1_hist.json.gz"), "wt") as f:
        json.dump(out, f)
    print("hist cases:", [(c["name"], [x["action"] for x in c["calls"]]) for c in cases])


def gen_uci_info():
    """`info depth` lines of the REAL player in UCI mode (uci=True, debugging=True; player.py:180-184,408-450) at
    search_threads=1 -> tests/golden/uci_info_k1.json.gz.  Wall-clock fields (time, nps) are dropped."""
    import contextlib
    import io
    import re
    import numpy as np
    from .ref_player_harness import FakeNetServer, make_config
    r = ref_import.senv()
    pm = ref_import.player_module()
    h31, h44 = _game_history(31, 41), _game_history(44, 42)
    specs = [
        dict(name="uci_init_320", seed=1, sims=320, state=r.INIT_STATE, turns=0, no_act=None, depth=None, hist=[r.INIT_STATE], use_history=False),
        dict(name="uci_black_31_depth3", seed=2, sims=800, state=h31[-1], turns=31, no_act=None, depth=300, hist=h31, use_history=False),
        dict(name="uci_red_44_no_act", seed=3, sims=250, state=h44[-1], turns=44, no_act="FIRST2", depth=None, hist=h44, use_history=False),
        dict(name="uci_black_31_history_planes", seed=4, sims=230, state=h31[-1], turns=31, no_act=None, depth=None, hist=h31, use_history=True),
    ]
    cases = []
    for sp in specs:
        cfg = make_config(sp["sims"], 1)
        srv = FakeNetServer()
        np.random.seed(sp["seed"])
        no_act = r.get_legal_moves(sp["state"])[:2] if sp["no_act"] == "FIRST2" else sp["no_act"]
        player = pm.CChessPlayer(cfg, pipes=srv.you, enable_resign=False, debugging=True, uci=True,
                                 use_history=sp["use_history"], side=sp["turns"] % 2)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            action, _ = player.action(sp["state"], sp["turns"], no_act=no_act, depth=sp["depth"], hist=list(sp["hist"]))
        node = player.tree[sp["state"]]
        lines = []
        for ln in buf.getvalue().splitlines():
            m = re.match(r"info depth (\d+) score (-?\d+) time \d+ pv(.*) nps -?\d+$", ln)
            assert m, ln
            lines.append([int(m.group(1)), int(m.group(2)), m.group(3)])
        value = float(player.debug[sp["state"]][1])
        cases.append({"name": sp["name"], "seed": sp["seed"], "sims": sp["sims"], "state": sp["state"], "turns": sp["turns"],
                      "no_act": no_act, "depth": sp["depth"], "hist": list(sp["hist"]), "use_history": sp["use_history"],
                      "action": action, "done_tasks": int(player.done_tasks), "root_value": value, "info": lines,
                      "legal": r.get_legal_moves(sp["state"]), "sum_n": int(node.sum_n),
                      "edges": {m: [int(x.n), float(x.w), float(x.q), float(x.p)] for m, x in node.a.items()}})
        player.close(wait=False)
        srv.close()
    notation = {"fen": "rnbakabnr/9/1c5c1/p1p1p1p1p/9/9/P1P1P1P1P/1C5C1/9/RNBAKABNR w - - 0 1",
                "fen_state": r.fen_to_state("rnbakabnr/9/1c5c1/p1p1p1p1p/9/9/P1P1P1P1P/1C5C1/9/RNBAKABNR w - - 0 1"),
                "ucci": [[m, r.parse_ucci_move(m)] for m in ("h2e2", "b9c7", "a0a1", "i9i0")],
                "uci": [[m, r.to_uci_move(m)] for m in ("7242", "1927", "0001", "8980")]}
    out = {"generator": "oracle/gen_golden_mcts.py:gen_uci_info", "reference": "NeymarL/ChineseChess-AlphaZero @7f45b0c agent/player.py",
           "cases": cases, "notation": notation}
    with gzip.open(os.path.join(GOLD, "uci_info_k1.json.gz"), "wt") as f:
        json.dump(out, f)
    print("uci cases:", [(c["name"], c["action"], c["info"]) for c in cases])


def gen_mcts_endgames():
    """K=1 searches of the real player on sparse endgames, where simulations run into terminal positions and IN-PATH
    repetitions (player.py:223-234: will_check_or_catch / be_catched verdicts) all the time -> mcts_k1_endgames.json.gz."""
    r = ref_import.senv()
    endings = ['3s5/9/9/9/4r4/9/9/4R4/9/4S4', '4s4/4m4/9/9/9/9/2R6/9/4M4/3S1r3', '2e1s4/4m4/4e4/9/9/9/9/4C4/4M4/3S5',
               '3s5/4m4/9/9/9/p8/9/4C4/4M4/4S4', '3s5/9/9/9/9/9/2k6/9/4R4/4S4', '5s3/9/9/2p6/9/9/6P2/9/9/3S5']
    endings = [s for s in endings if not r.done(s)[0]]
    assert len(endings) >= 5
    cases = []
    for i, s in enumerate(endings):
        for sims in (200, 500):
            res = real_player_moves([(s, 60 + i, None, False)], sims, 100 + i)
            a, e, sn = res[0]
            cases.append({"name": f"endgame{i}_{sims}", "seed": 100 + i, "sims": sims,
                          "calls": [{"state": s, "turns": 60 + i, "no_act": None, "increase_temp": False, "action": a, "sum_n": sn,
                                     "legal": r.get_legal_moves(s), "edges": {m: list(v) for m, v in e.items()}}]})
    out = {"generator": "oracle/gen_golden_mcts.py:gen_mcts_endgames", "reference": "NeymarL/ChineseChess-AlphaZero @7f45b0c agent/player.py",
           "config": {"search_threads": 1, "c_puct": 1.5, "noise_eps": 0.25, "dirichlet_alpha": 0.2, "tau_decay_rate": 0.98,
                      "virtual_loss": 3}, "cases": cases}
    with gzip.open(os.path.join(GOLD, "mcts_k1_endgames.json.gz"), "wt") as f:
        json.dump(out, f)
    print("endgame cases:", [(c["name"], c["calls"][0]["action"], c["calls"][0]["sum_n"]) for c in cases])


def gen_mcts_eps0():
    """noise_eps = 0: the reference still calls np.random.dirichlet once per legal move per root selection (player.py:304),
    so the position of the np.random stream after the search - and with it the sampled move - depends on it
    -> mcts_k1_eps0.json.gz (three plies of one game, moves sampled with tau > 0)."""
    import numpy as np
    from .ref_player_harness import FakeNetServer, make_config
    r = ref_import.senv()
    pm = ref_import.player_module()
    seed, sims = 31, 90
    cfg = make_config(sims, 1, noise_eps=0.0)
    srv = FakeNetServer()
    np.random.seed(seed)
    player = pm.CChessPlayer(cfg, pipes=srv.you, enable_resign=False)
    s, t, seq = r.INIT_STATE, 0, []
    for ply in range(4):
        a, _ = player.action(s, t)
        node = player.tree[s]
        seq.append({"state": s, "turns": t, "no_act": None, "increase_temp": False, "action": a, "sum_n": int(node.sum_n),
                    "legal": r.get_legal_moves(s),
                    "edges": {m: [int(x.n), float(x.w), float(x.q), float(x.p)] for m, x in node.a.items()}})
        s = r.step(s, a)
        t += 1
    tail = float(np.random.rand())
    player.close(wait=False)
    srv.close()
    out = {"generator": "oracle/gen_golden_mcts.py:gen_mcts_eps0", "reference": "NeymarL/ChineseChess-AlphaZero @7f45b0c agent/player.py",
           "config": {"search_threads": 1, "c_puct": 1.5, "noise_eps": 0.0, "dirichlet_alpha": 0.2, "tau_decay_rate": 0.98, "virtual_loss": 3},
           "cases": [{"name": "eps0_four_plies", "seed": seed, "sims": sims, "calls": seq, "rand_after": tail}]}
    with gzip.open(os.path.join(GOLD, "mcts_k1_eps0.json.gz"), "wt") as f:
        json.dump(out, f)
    print("eps0:", [c["action"] for c in seq], tail)
