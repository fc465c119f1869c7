"""Large random-playout sweep of the rules kernels against the oracle (>= 1e5 positions).

Positions come from seeded random playouts of oracle/senv.py (itself pinned to the real reference on 3 190 + 2 145
positions, tests/test_oracle_vs_reference.py / tests/golden/env_playouts.json.gz); the oracle's answers are packed into
numpy arrays by worker processes (the oracle is ~0.5 ms per position, so 1e5 positions take a few seconds on the GPU box's
host cores) and compared with the kernels' answers in bulk: ordered move lists, done / final_move / check, planes,
step + flip, no_eat, will_check_or_catch, be_catched, has_attack_chessman — bit for bit."""
import multiprocessing as mp
import os
import random

import numpy as np


def _u16(m):
    return ((int(m[1]) * 9 + int(m[0])) << 8) | (int(m[3]) * 9 + int(m[2]))


def _playout_rows(args):
    seed, n_rows = args
    from oracle import senv as o
    rng = random.Random(seed)
    states, nexts = [], []
    moves = np.full((n_rows, 128), 0xFFFF, dtype=np.uint16)
    counts = np.zeros(n_rows, dtype=np.int32)
    done = np.zeros((n_rows, 3), dtype=np.int8)
    final = np.full(n_rows, 0xFFFF, dtype=np.uint16)
    planes = np.zeros((n_rows, 158), dtype=np.uint8)
    move = np.full(n_rows, 0xFFFF, dtype=np.uint16)
    flags = np.zeros((n_rows, 4), dtype=np.uint8)          # no_eat, wcc, bc, attack
    i = 0
    while i < n_rows:
        s = o.INIT_STATE
        for _ in range(300):
            if i >= n_rows:
                break
            lm = o.get_legal_moves(s)
            d = o.done(s, need_check=True)
            states.append(s)
            counts[i] = len(lm)
            moves[i, :len(lm)] = [_u16(m) for m in lm]
            done[i] = (d[0], d[1], d[3] if len(d) == 4 else 0)
            if d[2] is not None:
                final[i] = _u16(d[2])
            planes[i] = np.packbits(o.state_to_planes(s).reshape(-1).astype(np.uint8))[:158]
            flags[i, 3] = o.has_attack_chessman(s)
            if d[0] or not lm:
                nexts.append(None)
                i += 1
                break
            m = rng.choice(lm)
            ns, ne = o.new_step(s, m)
            move[i] = _u16(m)
            flags[i, 0] = ne
            flags[i, 1] = o.will_check_or_catch(s, m)
            flags[i, 2] = o.be_catched(s, m)
            nexts.append(ns)
            s = ns
            i += 1
    return states, nexts, moves, counts, done, final, planes, move, flags


def oracle_rows(n_positions, seed, procs=None):
    procs = procs or max(1, min(48, (os.cpu_count() or 2) - 1))
    per = 2500
    jobs = [(seed * 100003 + k, min(per, n_positions - k * per)) for k in range((n_positions + per - 1) // per)]
    if procs == 1 or len(jobs) == 1:
        parts = [_playout_rows(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(min(procs, len(jobs))) as pool:
            parts = pool.map(_playout_rows, jobs)
    states = [s for p in parts for s in p[0]]
    nexts = [s for p in parts for s in p[1]]
    arrs = [np.concatenate([p[k] for p in parts]) for k in range(2, 9)]
    return (states, nexts) + tuple(arrs)


def check_sweep(env, n_positions, seed=2024, procs=None, chunk=20000):
    from cczero_b200.env import board_to_state
    states, nexts, moves, counts, done, final, planes, move, flags = oracle_rows(n_positions, seed, procs)
    n = len(states)
    assert n == n_positions
    stats = {"positions": n, "moves": int(counts.sum()), "max_moves": int(counts.max()), "terminal": int(done[:, 0].sum()),
             "checks": int(done[:, 2].sum()), "wcc": int(flags[:, 1].sum()), "bc": int(flags[:, 2].sum()),
             "captures": int((flags[:, 0] == 0)[move != 0xFFFF].sum())}
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        boards = env.boards_from_states(states[a:b])
        mv, cnt = env.movegen_batch(boards)
        assert np.array_equal(cnt.cpu().numpy(), counts[a:b])
        assert np.array_equal(mv.cpu().numpy().view(np.uint16), moves[a:b])          # ordered lists AND the 0xFFFF padding
        out, fm = env.done_batch(boards, need_check=True)
        out, fm = out.cpu().numpy(), fm.cpu().numpy().view(np.uint16)
        assert np.array_equal(out[:, 0], done[a:b, 0]) and np.array_equal(out[:, 1], done[a:b, 1])
        live = done[a:b, 0] == 0                                                     # the reference only computes `check` then
        assert np.array_equal(out[live, 2], done[a:b, 2][live])
        assert np.array_equal(fm, final[a:b])
        pl = env.planes_batch(boards).cpu().numpy().reshape(b - a, -1)
        assert set(np.unique(pl)) <= {0.0, 1.0}
        assert np.array_equal(np.packbits(pl.astype(np.uint8), axis=1)[:, :158], planes[a:b])
        idx = np.nonzero(move[a:b] != 0xFFFF)[0]
        sub = boards[idx]
        import torch
        mt = torch.as_tensor(move[a:b][idx].view(np.int16)).to(boards.device)
        nb, ne = env.step_batch(sub, mt)
        nb = nb.cpu().numpy()
        assert np.array_equal(ne.cpu().numpy().astype(np.uint8), flags[a:b, 0][idx])
        want = env.boards_from_states([nexts[a + k] for k in idx]).cpu().numpy()
        assert np.array_equal(nb, want)
        for j in range(0, len(idx), 997):                                            # string codec on a sample
            assert board_to_state(nb[j]) == nexts[a + idx[j]]
        wcc, bc, ha = env.check_catch_batch(sub, mt)
        assert np.array_equal(wcc.cpu().numpy().astype(np.uint8), flags[a:b, 1][idx])
        assert np.array_equal(bc.cpu().numpy().astype(np.uint8), flags[a:b, 2][idx])
        _, _, ha_all = env.check_catch_batch(boards, torch.zeros(b - a, dtype=torch.int16, device=boards.device))
        assert np.array_equal(ha_all.cpu().numpy().astype(np.uint8), flags[a:b, 3])
    return stats
