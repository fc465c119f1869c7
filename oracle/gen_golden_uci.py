"""A scripted session with the REAL reference UCI front end (cchess_alphazero/uci.py, class UCI, unmodified; Keras /
TensorFlow imports stubbed, the model replaced by the deterministic fake network behind real Pipes) at
search_threads = 1 -> tests/golden/uci_session_k1.json.gz.  Build-container only (python -m oracle.gen_golden uci_session).
Wall-clock fields (`time`, `nps`) are stripped from the recorded lines."""
import contextlib
import gzip
import io
import json
import os
import re
import sys
import time

import numpy as np

from . import ref_worker_harness as h
from .ref_player_harness import FakeNetServer

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

PLAY = dict(search_threads=1, simulation_num_per_move=100, c_puct=1.5, noise_eps=0.2, dirichlet_alpha=0.2, tau_decay_rate=0.9,
            virtual_loss=3, resign_threshold=-0.99, min_resign_turn=40, max_game_length=200)

# (np.random seed or None, command); `go` commands are followed until `bestmove`
SCRIPT = [
    (None, "uci"), (None, "isready"),
    (None, "position startpos moves h2e2 h9g7"), (5, "go depth 2"),
    (None, "position fen rnbakabnr/9/1c5c1/p1p1p1p1p/9/9/P1P1P1P1P/1C2C4/9/RNBAKABNR b - - 0 1"), (6, "go depth 1"),
    (None, "fen rnbakab1r/9/1c4nc1/p1p1p1p1p/9/9/P1P1P1P1P/1C2C4/9/RNBAKABNR w - - 0 2 moves b0c2 b9c7"), (7, "go depth 1"),
    (None, "position startpos moves b0c2 b9c7 c2b0 c7b9"), (8, "go depth 1"),
    (None, "position startpos moves b0c2 b9c7 c2b0 c7b9 b0c2 b9c7 c2b0"), (9, "go depth 1"),
    (None, "ucinewgame"), (None, "position startpos"), (10, "go depth 1"),
    (None, "position moves a0a1"), (11, "go depth 1"),
]


def strip_clock(line):
    line = re.sub(r" time \d+", "", line)
    return re.sub(r" nps -?\d+", "", line)


def gen_uci_session():
    h.worker_modules()
    err = sys.stderr
    import cchess_alphazero.uci as ruci            # module level: builds a Config, redirects stderr to a log file
    sys.stderr = err
    cfg = ruci.config
    for k, v in PLAY.items():
        setattr(cfg.play, k, v)
    servers = []

    class FakeModel:
        def get_pipes(self, need_reload=True):
            servers.append(FakeNetServer())
            return servers[-1].you

        def close_pipes(self):
            pass
    u = ruci.UCI(cfg)
    u.load_model = lambda config_file=None: (setattr(u, "model", FakeModel()) or False)
    ruci.set_session_config = lambda **k: None
    steps = []
    for seed, cmd in SCRIPT:
        buf = io.StringIO()
        parts = cmd.split(" ")
        u.args = parts[1:]
        if seed is not None:
            np.random.seed(seed)
        with contextlib.redirect_stdout(buf):
            getattr(u, "cmd_" + parts[0])()
            if parts[0] == "go":
                t0 = time.time()
                while "bestmove" not in buf.getvalue() and time.time() - t0 < 120:
                    time.sleep(0.02)
                time.sleep(0.1)
        steps.append({"seed": seed, "cmd": cmd, "out": [strip_clock(x) for x in buf.getvalue().splitlines()],
                      "state": u.state, "turns": u.turns, "is_red_turn": u.is_red_turn})
    for s in servers:
        s.close()
    out = {"generator": "oracle/gen_golden_uci.py", "reference": "NeymarL/ChineseChess-AlphaZero @7f45b0c uci.py", "play": PLAY,
           "steps": steps}
    with gzip.open(os.path.join(GOLD, "uci_session_k1.json.gz"), "wt") as f:
        json.dump(out, f)
    for s in steps:
        print(s["cmd"], "->", s["out"][-2:] if s["out"] else "")
