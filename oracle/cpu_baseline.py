"""CPU baseline worker: the reference's self-play hot path (oracle port) on ONE host core.  TEST/BENCH INFRASTRUCTURE.

The reference scales on CPU with processes (`config.play.max_processes`, worker/self_play.py:55-60), each a
single-threaded Python player; torch intra-op threads on these tiny batches only oversubscribe (SURVEY.md §8d /
BASELINE.md §4).  bench.py therefore runs one of these workers per host core and sums their rates.
"""
import os
import time


def worker(args):
    filters, blocks, sims, k, budget_s, seed = args
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    import numpy as np
    import torch
    torch.set_num_threads(1)
    from oracle import model as om
    from oracle import player as op
    from oracle import senv
    w = om.init_weights(filters, blocks, 256, seed=0)
    net = om.TorchNet(w, blocks)
    deadline = [0.0]

    class Stop(Exception):
        pass

    def evaluate(states):
        if time.time() > deadline[0]:
            raise Stop()
        planes = np.stack([senv.state_to_planes(s) for s in states])
        p, v = net.predict_on_batch(planes)
        return [(p[i], float(v[i, 0])) for i in range(len(states))]

    pc = op.PlayConfig(simulation_num_per_move=sims, search_threads=k, c_puct=1.5, noise_eps=0.15, dirichlet_alpha=0.2,
                       tau_decay_rate=0.9, virtual_loss=3)
    np.random.seed(seed)
    pl = op.OraclePlayer(pc, evaluate)
    evaluate([senv.INIT_STATE])                 # warm the torch kernels outside the window
    t0 = time.time()
    deadline[0] = t0 + budget_s
    state, turns = senv.INIT_STATE, 0
    try:
        while True:                             # keep playing moves until the window closes
            a, _ = pl.action(state, turns)
            if a is None:
                break
            state = senv.step(state, a)
            turns += 1
            if senv.done(state)[0]:
                state, turns = senv.INIT_STATE, 0
                pl = op.OraclePlayer(pc, evaluate)
    except Stop:
        pass
    dt = time.time() - t0
    return pl.stats["positions"] if turns == 0 else None, dt, pl.stats


def run(filters, blocks, sims, k, budget_s, n_procs):
    """Returns (aggregate sims/s, total sims, mean window seconds, n_procs).  One simulation ~ one evaluated position."""
    import multiprocessing as mp
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    ctx = mp.get_context("spawn")
    with ctx.Pool(n_procs) as pool:
        res = pool.map(_count_worker, [(filters, blocks, sims, k, budget_s, i) for i in range(n_procs)])
    rate = sum(n / dt for n, dt in res if dt > 0)
    return rate, sum(n for n, _ in res), sum(dt for _, dt in res) / len(res), n_procs


def _count_worker(args):
    """Simulations completed in the window = tasks finished (leaf evaluations + terminal hits), counted by the player."""
    filters, blocks, sims, k, budget_s, seed = args
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    import numpy as np
    import torch
    torch.set_num_threads(1)
    from oracle import model as om
    from oracle import player as op
    from oracle import senv
    net = om.TorchNet(om.init_weights(filters, blocks, 256, seed=0), blocks)
    deadline = [float("inf")]
    done = [0]

    class Stop(Exception):
        pass

    def evaluate(states):
        if time.time() > deadline[0]:
            raise Stop()
        planes = np.stack([senv.state_to_planes(s) for s in states])
        p, v = net.predict_on_batch(planes)
        done[0] += len(states)
        return [(p[i], float(v[i, 0])) for i in range(len(states))]

    pc = op.PlayConfig(simulation_num_per_move=sims, search_threads=k, c_puct=1.5, noise_eps=0.15, dirichlet_alpha=0.2,
                       tau_decay_rate=0.9, virtual_loss=3)
    np.random.seed(seed)
    evaluate([senv.INIT_STATE])
    done[0] = 0
    t0 = time.time()
    deadline[0] = t0 + budget_s
    state, turns = senv.INIT_STATE, 0
    pl = op.OraclePlayer(pc, evaluate)
    try:
        while True:
            a, _ = pl.action(state, turns)
            state = senv.step(state, a)
            turns += 1
            if senv.done(state)[0] or turns >= 200:
                state, turns = senv.INIT_STATE, 0
                pl = op.OraclePlayer(pc, evaluate)
    except Stop:
        pass
    return done[0], time.time() - t0
