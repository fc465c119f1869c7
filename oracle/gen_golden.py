"""Generate tests/golden/*.json.gz from the REAL reference (needs /root/reference; run in the build
container only).  The fixtures pin both the oracle restatements and the CUDA kernels.

  python oracle/gen_golden.py env      # rules-engine vectors from random playouts
  python oracle/gen_golden.py mcts     # K=1 seeded searches of the real CChessPlayer (fake NN)
"""
import gzip
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_import  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def gen_env(n_games=40, seed=20240922):
    r = ref_import.senv()
    rng = random.Random(seed)
    rows = []
    for g in range(n_games):
        s = r.INIT_STATE
        for ply in range(160):
            lm = r.get_legal_moves(s)
            d = r.done(s, need_check=True)
            row = {"state": s, "moves": lm, "done": list(d), "attack": r.has_attack_chessman(s),
                   "flip": r.fliped_state(s),
                   "plane_idx": [int(i) for i in r.state_to_planes(s).reshape(-1).nonzero()[0]]}
            if d[0] or not lm:
                rows.append(row)
                break
            m = rng.choice(lm)
            ns, no_eat = r.new_step(s, m)
            row.update({"move": m, "next": ns, "no_eat": no_eat,
                        "wcc": bool(r.will_check_or_catch(s, m)), "bc": bool(r.be_catched(s, m))})
            rows.append(row)
            s = ns
    extra = ['4s4/9/4e4/p8/2e2R2p/P5E2/8P/9/9/4S1E2', '4s4/9/9/9/9/9/9/9/9/4S4',
             'rkemsmek1/8r/1c5c1/p1p1p1p1p/9/9/P1P1P1P1P/1C5C1/9/RKEMSMEKR']
    for s in extra:
        rows.append({"state": s, "moves": r.get_legal_moves(s), "done": list(r.done(s, need_check=True)),
                     "attack": r.has_attack_chessman(s), "flip": r.fliped_state(s),
                     "plane_idx": [int(i) for i in r.state_to_planes(s).reshape(-1).nonzero()[0]]})
    lt = ref_import.lookup_tables()
    out = {"generator": "oracle/gen_golden.py env", "reference": "NeymarL/ChineseChess-AlphaZero @7f45b0c",
           "labels_first": lt.ActionLabelsRed[:12], "labels_last": lt.ActionLabelsRed[-5:],
           "n_labels": len(lt.ActionLabelsRed), "labels_sha": __import__("hashlib").sha256("".join(lt.ActionLabelsRed).encode()).hexdigest(),
           "rows": rows}
    os.makedirs(GOLD, exist_ok=True)
    with gzip.open(os.path.join(GOLD, "env_playouts.json.gz"), "wt") as f:
        json.dump(out, f)
    print("env rows:", len(rows))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "env"
    if what == "env":
        gen_env()
    elif what == "mcts":
        from oracle.gen_golden_mcts import gen_mcts
        gen_mcts()
    elif what == "uci_session":
        from oracle.gen_golden_uci import gen_uci_session
        gen_uci_session()
    elif what == "games":
        from oracle.gen_golden_games import gen_games
        gen_games()
    elif what == "eps0":
        from oracle.gen_golden_mcts import gen_mcts_eps0
        gen_mcts_eps0()
    elif what == "endgames":
        from oracle.gen_golden_mcts import gen_mcts_endgames
        gen_mcts_endgames()
    elif what == "uci":
        from oracle.gen_golden_mcts import gen_uci_info
        gen_uci_info()
    elif what == "mcts_hist":
        from oracle.gen_golden_mcts import gen_mcts_hist
        gen_mcts_hist()
