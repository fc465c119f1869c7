"""Does running the residual tower in L2-sized chunks beat one big batch? (development experiment)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cczero_b200.engine import Engine
from cczero_b200.env import state_to_board
from cczero_b200.lib import get_lib
from oracle import model as om, senv
from tools.bench_nn import time_it

lib = get_lib()
w = {k: torch.as_tensor(v) for k, v in om.init_weights(256, 20, 256, seed=0).items()}
boards = torch.zeros(8192, 96, dtype=torch.uint8, device="cuda")
boards[:] = torch.as_tensor(state_to_board(senv.INIT_STATE)).cuda()
for skip in (False, True):
    for chunk in (210, 421, 631, 1052, 2104, 8192):
        eng = Engine(lib, "cuda", n_games=chunk, sims_per_move=8, leaves_per_round=1, nn_filters=256, nn_blocks=20, nn_fp32_skip=skip)
        eng.set_weights(w)
        ms = time_it(lambda: eng.nn_forward_boards(boards), iters=4, warm=2)
        print(f"fp32_skip={skip} chunk={chunk}: {ms:.2f} ms  {8192 / ms * 1e3:.0f} pos/s", flush=True)
        eng.close()
