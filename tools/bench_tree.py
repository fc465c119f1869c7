"""Integer kernels alone: the search with a constant evaluator (uniform policy, zero value) so that only k_wave / k_scan /
k_gather / k_apply and the host round trip are timed.  python tools/bench_tree.py [games] [sims] [K]"""
import sys
import time

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from cczero_b200.engine import Engine   # noqa: E402
from cczero_b200.lib import get_lib     # noqa: E402

games = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 800
K = int(sys.argv[3]) if len(sys.argv) > 3 else 8
eng = Engine(get_lib(), "cuda", n_games=games, sims_per_move=sims, leaves_per_round=K, noise_mode=1, noise_eps=0.25)
eng.reset()
pol = torch.full((games * K, 2086), 1.0 / 2086, dtype=torch.float32, device="cuda")
val = torch.zeros(games * K, dtype=torch.float32, device="cuda")
for move in range(3):
    eng.search_begin(None)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    waves = leaves = 0
    while True:
        n, busy = eng.search_wave()
        waves += 1
        leaves += n
        if n:
            eng.search_apply(pol, val)
        if not busy:
            break
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = eng.search_stats()
    print(f"move {move}: {waves} waves, {leaves} leaves, {dt * 1e3:.1f} ms, {dt / waves * 1e6:.0f} us per wave+apply, "
          f"{games * sims / dt / 1e6:.2f} M sims/s (tree kernels + host sync only); mean path {st['path_edges'] / max(1, st['sims']):.2f}")
    eng.play_move()
eng.close()
