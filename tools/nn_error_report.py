"""Max |error| of the tensor-core forward against the fp32 restatement for the nets of BASELINE.json."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cczero_b200.engine import Engine
from cczero_b200.env import StaticEnv
from cczero_b200.lib import get_lib
from oracle import model as om, senv
from tests.search_checks import midgame_states

lib = get_lib()
env = StaticEnv(lib, "cuda")
states = [senv.INIT_STATE] + midgame_states(63, 3, lo=1, hi=120)
planes = np.stack([senv.state_to_planes(s) for s in states])
import sys as _s
mode = {"fp16": False, "fp32": True, "auto": None}[_s.argv[1] if len(_s.argv) > 1 else "auto"]
for f, b, trained in ((128, 7, False), (128, 7, True), (256, 7, False), (192, 10, True), (256, 20, False), (256, 20, True)):
    w = om.init_weights(f, b, 256, seed=1, trained_like=trained)
    rp, rv = om.forward(w, planes, b)
    eng = Engine(lib, "cuda", n_games=64, sims_per_move=8, leaves_per_round=1, nn_filters=f, nn_blocks=b, nn_fp32_skip=mode)
    eng.set_weights({k: torch.as_tensor(v) for k, v in w.items()})
    p, v = eng.nn_forward_boards(env.boards_from_states(states))
    p, v = p.cpu().numpy(), v.cpu().numpy()
    print(f"{f}x{b} trained_like={trained}: max|dp|={np.abs(p-rp).max():.2e} max|dv|={np.abs(v-rv).max():.2e} "
          f"max p={rp.max():.3f} |v| range=({np.abs(rv).min():.3f},{np.abs(rv).max():.3f}) argmax agree={(p.argmax(1)==rp.argmax(1)).mean():.2f}")
    eng.close()
