"""UCI-mode features of the player (reference: agent/player.py:180-184,408-450 `info depth` lines, :88-106
close_and_return_action, `depth` / `infinite`) and the front end cczero_b200/uci.py (reference: uci.py:40-331).
Pinned by tests/golden/uci_info_k1.json.gz: the lines the REAL player printed at search_threads = 1."""
import gzip
import io
import json
import os
import threading
import time

import numpy as np
import pytest

from oracle import player as op
from oracle import senv as osenv
from tests import search_checks as sc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gold():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "uci_info_k1.json.gz"), "rt") as f:
        return json.load(f)


def test_notation_helpers_match_reference():
    n = _gold()["notation"]
    assert osenv.fen_to_state(n["fen"]) == n["fen_state"] == osenv.INIT_STATE
    for a, b in n["ucci"]:
        assert osenv.parse_ucci_move(a) == b
    for a, b in n["uci"]:
        assert osenv.to_uci_move(a) == b
    from cczero_b200 import env as penv
    assert penv.fen_to_state(n["fen"]) == n["fen_state"]
    assert [penv.parse_ucci_move(a) for a, _ in n["ucci"]] == [b for _, b in n["ucci"]]
    assert [penv.to_uci_move(a) for a, _ in n["uci"]] == [b for _, b in n["uci"]]


def test_oracle_info_lines_match_real_player():
    for case in _gold()["cases"]:
        pc = op.PlayConfig(simulation_num_per_move=case["sims"], search_threads=1, c_puct=1.5, noise_eps=0.25,
                           dirichlet_alpha=0.2, tau_decay_rate=0.98, virtual_loss=3, resign_threshold=-0.92, min_resign_turn=20)
        np.random.seed(case["seed"])
        ev = op.fake_evaluate_states_hist if case["use_history"] else op.fake_evaluate_states
        pl = op.OraclePlayer(pc, ev, use_history=case["use_history"], uci=True, debugging=True, side=case["turns"] % 2)
        a, _ = pl.action(case["state"], case["turns"], case["no_act"], depth=case["depth"], hist=case["hist"])
        assert [list(x) for x in pl.info] == case["info"], case["name"]
        assert a == case["action"] and pl.done_tasks == case["done_tasks"]
        assert float(pl.debug[case["state"]][1]) == case["root_value"]


def check_player_uci(lib, device):
    """The drop-in player in UCI mode: same `info depth` lines (depth, score, pv), same move, same root statistics."""
    from cczero_b200.player import CChessPlayer
    for case in _gold()["cases"]:
        srv = sc.FakeNetServer()
        np.random.seed(case["seed"])
        player = CChessPlayer(sc.make_config(case["sims"], 1), pipes=srv.you, lib=lib, device=device, debugging=True, uci=True,
                              use_history=case["use_history"], side=case["turns"] % 2)
        out = io.StringIO()
        player.info_stream = out
        try:
            action, _ = player.action(case["state"], case["turns"], no_act=case["no_act"], depth=case["depth"], hist=case["hist"])
            lines = []
            for ln in out.getvalue().splitlines():
                parts = ln.split(" ")
                assert parts[:2] == ["info", "depth"] and parts[3] == "score" and parts[5] == "time" and parts[7] == "pv"
                lines.append([int(parts[2]), int(parts[4]), "".join(" " + m for m in parts[8:parts.index("nps")])])
            assert lines == case["info"], (case["name"], lines, case["info"])
            assert action == case["action"] and player.done_tasks == case["done_tasks"]
            assert float(np.float32(player.debug[case["state"]][1])) == float(np.float32(case["root_value"]))
            root = player.engine.root(0)
            assert root["sum_n"] == case["sum_n"]
            for m, n, w in zip(root["moves"], root["n"], root["w"]):
                gn, gw = case["edges"].get(m, [0, 0.0])[:2]
                assert (n, w) == (gn, gw)
        finally:
            player.close()
            srv.close()


def check_infinite_and_stop(lib, device):
    """`go infinite` + `stop` (uci.py:229-243, player.py:88-106): the search runs until close_and_return_action, which
    answers from the tree as it stands."""
    from cczero_b200.player import CChessPlayer
    srv = sc.FakeNetServer()
    np.random.seed(3)
    player = CChessPlayer(sc.make_config(100, 8), pipes=srv.you, lib=lib, device=device, debugging=True, uci=True,
                          infinite_capacity=6000)
    player.info_stream = io.StringIO()
    res = {}
    th = threading.Thread(target=lambda: res.update(r=player.action(osenv.INIT_STATE, 0, infinite=True)), daemon=True)
    th.start()
    t0 = time.time()
    while player.done_tasks < 300 and time.time() - t0 < 120:
        time.sleep(0.01)
    got = player.close_and_return_action(osenv.INIT_STATE, 0, None)
    th.join(60)
    assert not th.is_alive()
    action, value, depth = got
    assert action in osenv.get_legal_moves(osenv.INIT_STATE) and depth >= 3 and -1 <= value <= 1
    assert 300 <= player.done_tasks < 100000
    player.close()
    srv.close()


class _Lines:
    """stdout stand-in that lets the test wait for a line."""

    def __init__(self):
        self.lines, self.cv = [], threading.Condition()

    def write(self, text):
        with self.cv:
            for ln in text.splitlines():
                if ln:
                    self.lines.append(ln)
            self.cv.notify_all()

    def flush(self):
        pass

    def wait_for(self, prefix, start=0, timeout=120):
        t0 = time.time()
        with self.cv:
            while True:
                for i in range(start, len(self.lines)):
                    if self.lines[i].startswith(prefix):
                        return i
                if time.time() - t0 > timeout:
                    raise AssertionError(f"no line starting with {prefix!r}: {self.lines[start:]}")
                self.cv.wait(0.5)


def check_uci_session(lib, device):
    """A scripted session against cczero_b200/uci.py (command set and answers of the reference's uci.py:59-331)."""
    from cczero_b200.uci import UCI
    from types import SimpleNamespace
    servers = []

    def pipes_factory():
        srv = sc.FakeNetServer()
        servers.append(srv)
        return srv.you
    out = _Lines()
    cfg = sc.make_config(800, 10)
    u = UCI(cfg, model=SimpleNamespace(use_history=False), lib=lib, device=device, stdout=out, pipes_factory=pipes_factory,
            infinite_capacity=5000)
    np.random.seed(11)

    def send(cmd):
        parts = cmd.split(' ')
        u.args = parts[1:]
        return getattr(u, 'cmd_' + parts[0])()
    send("uci")
    assert out.lines[:6] == ['id name CCZero', 'id author https://cczero.org', 'id version 2.4',
                             'option name gpu spin default 0 min 0 max 7', 'option name Threads spin default 10 min 0 max 1024', 'uciok']
    send("isready")
    assert out.lines[-1] == 'readyok'
    send("setoption name Threads value 8")
    assert cfg.play.search_threads == 8
    # position + moves: red h2e2 (7242), black h9g7 (from black's side of the board)
    send("position startpos moves h2e2 h9g7")
    s1 = osenv.step(osenv.INIT_STATE, "7242")
    s2 = osenv.step(s1, osenv.flip_move("7967"))
    assert u.state == s2 and u.turns == 2 and u.is_red_turn and u.history == [osenv.INIT_STATE, "7242", s1, osenv.flip_move("7967"), s2]
    # go depth 2 = 200 simulations: two info lines with a pv, then the summary and the best move
    n0 = len(out.lines)
    send("go depth 2")
    i = out.wait_for("bestmove", n0)
    infos = [ln for ln in out.lines[n0:i] if ln.startswith("info depth")]
    assert len(infos) == 3 and " pv " in infos[0] and infos[0].startswith("info depth 1 ") and infos[1].startswith("info depth 2 ")
    assert " pv " not in infos[2] and infos[2].startswith("info depth 2 score ")
    best = out.lines[i].split(' ')
    assert osenv.parse_ucci_move(best[1]) in osenv.get_legal_moves(s2)
    if len(best) > 2:                                        # ponder: a legal reply, written from black's side
        assert best[2] == "ponder"
        s3 = osenv.step(s2, osenv.parse_ucci_move(best[1]))
        assert osenv.flip_move(osenv.parse_ucci_move(best[3])) in osenv.get_legal_moves(s3)
    u.search_worker.join(30)
    # black to move from a FEN; infinite search stopped by `stop`
    send("position fen rnbakabnr/9/1c5c1/p1p1p1p1p/9/9/P1P1P1P1P/1C2C4/9/RNBAKABNR b - - 0 1")
    assert not u.is_red_turn and u.turns == 1 and u.state == osenv.step(osenv.INIT_STATE, "7242")
    n0 = len(out.lines)
    send("go infinite")
    out.wait_for("info depth 2", n0)
    send("stop")
    i = out.wait_for("bestmove", n0)
    mv = osenv.flip_move(osenv.parse_ucci_move(out.lines[i].split(' ')[1]))     # black's move is printed in board coordinates
    assert mv in osenv.get_legal_moves(u.state)
    u.search_worker.join(30)
    assert not u.search_worker.is_alive()
    # movetime: the timer stops the search
    n0 = len(out.lines)
    send("go movetime 1500")
    out.wait_for("bestmove", n0, timeout=60)
    u.search_worker.join(30)
    assert send("quit") == "quit"
    for srv in servers:
        srv.close()


def check_uci_session_against_real_front_end(lib, device):
    """tests/golden/uci_session_k1.json.gz: a session with the REAL reference uci.UCI (search_threads = 1, fake network).
    cczero_b200/uci.py must print the same lines (clock fields stripped) and hold the same position after every command."""
    import re
    from types import SimpleNamespace
    from cczero_b200.uci import UCI
    with gzip.open(os.path.join(ROOT, "tests", "golden", "uci_session_k1.json.gz"), "rt") as f:
        gold = json.load(f)
    p = gold["play"]
    cfg = sc.make_config(p["simulation_num_per_move"], p["search_threads"], **{k: v for k, v in p.items()
                                                                                if k not in ("simulation_num_per_move", "search_threads")})
    servers = []

    def pipes_factory():
        servers.append(sc.FakeNetServer())
        return servers[-1].you
    out = _Lines()
    u = UCI(cfg, model=SimpleNamespace(use_history=False), lib=lib, device=device, stdout=out, pipes_factory=pipes_factory,
            infinite_capacity=4000)
    for step in gold["steps"]:
        n0 = len(out.lines)
        parts = step["cmd"].split(' ')
        u.args = parts[1:]
        if step["seed"] is not None:
            np.random.seed(step["seed"])
        getattr(u, 'cmd_' + parts[0])()
        if parts[0] == "go":
            out.wait_for("bestmove", n0)
            u.search_worker.join(30)
        got = [re.sub(r" nps -?\d+", "", re.sub(r" time \d+", "", ln)) for ln in out.lines[n0:]]
        assert got == step["out"], (step["cmd"], got, step["out"])
        assert (u.state, u.turns, u.is_red_turn) == (step["state"], step["turns"], step["is_red_turn"]), step["cmd"]
    for srv in servers:
        srv.close()


def test_emul_uci_session_equals_real_front_end(emul_lib):
    check_uci_session_against_real_front_end(emul_lib, "cpu")


def test_emul_uci_session(emul_lib):
    check_uci_session(emul_lib, "cpu")


def test_emul_player_uci(emul_lib):
    check_player_uci(emul_lib, "cpu")


def test_emul_infinite_and_stop(emul_lib):
    check_infinite_and_stop(emul_lib, "cpu")


@pytest.mark.gpu
def test_cuda_player_uci(cuda_lib):
    check_player_uci(cuda_lib, "cuda")
    check_infinite_and_stop(cuda_lib, "cuda")
    check_uci_session(cuda_lib, "cuda")
    check_uci_session_against_real_front_end(cuda_lib, "cuda")
