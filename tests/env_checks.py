"""Shared parity checks for the rules kernels: the same assertions run against the CUDA library
(-m gpu) and against the CPU SIMT-emulation build of the same kernel source (-m "not gpu")."""
import numpy as np

from cczero_b200.env import (board_to_state, move_to_u16, state_to_board, u16_to_move)
from oracle import senv as osenv


def check_against_rows(env, rows):
    """rows: golden dicts produced by the real reference (oracle/gen_golden.py)."""
    states = [r["state"] for r in rows]
    boards = env.boards_from_states(states)
    n = len(rows)
    # codec round trip
    for s in states[:50]:
        assert board_to_state(state_to_board(s)) == s
    # movegen: ordered list equality
    mv, cnt = env.movegen_batch(boards)
    mv = mv.cpu().numpy().view(np.uint16)
    cnt = cnt.cpu().numpy()
    for i, r in enumerate(rows):
        got = [u16_to_move(v) for v in mv[i, :cnt[i]]]
        assert got == r["moves"], (r["state"], got, r["moves"])
        assert (mv[i, cnt[i]:] == 0xFFFF).all()
    # done(need_check=True)
    out, fm = env.done_batch(boards, need_check=True)
    out = out.cpu().numpy()
    fm = fm.cpu().numpy().view(np.uint16)
    for i, r in enumerate(rows):
        d = r["done"]
        assert bool(out[i, 0]) == d[0] and int(out[i, 1]) == d[1], (r["state"], out[i], d)
        assert (None if fm[i] == 0xFFFF else u16_to_move(fm[i])) == d[2], (r["state"], fm[i], d)
        if len(d) == 4:
            assert bool(out[i, 2]) == d[3], (r["state"], out[i], d)
    # planes
    planes = env.planes_batch(boards).cpu().numpy().reshape(n, -1)
    for i, r in enumerate(rows):
        assert planes[i].nonzero()[0].tolist() == r["plane_idx"], r["state"]
        assert set(np.unique(planes[i])) <= {0.0, 1.0}
    # flip
    for r in rows[:200]:
        assert env.fliped_state(r["state"]) == r["flip"]
    # step / new_step, check & catch on the rows that carry a move
    idx = [i for i, r in enumerate(rows) if "move" in r]
    sub = boards[idx]
    moves = env.moves_tensor([rows[i]["move"] for i in idx])
    nb, ne = env.step_batch(sub, moves)
    nb = nb.cpu().numpy()
    ne = ne.cpu().numpy()
    wcc, bc, ha = env.check_catch_batch(sub, moves)
    wcc, bc, ha = wcc.cpu().numpy(), bc.cpu().numpy(), ha.cpu().numpy()
    for k, i in enumerate(idx):
        r = rows[i]
        assert board_to_state(nb[k]) == r["next"], r["state"]
        assert (nb[k, 90:] == 0).all()
        assert bool(ne[k]) == r["no_eat"]
        assert bool(wcc[k]) == r["wcc"], (r["state"], r["move"])
        assert bool(bc[k]) == r["bc"], (r["state"], r["move"])
        assert bool(ha[k]) == r["attack"]


def check_keys(env, rows):
    """Position keys: equal boards <=> equal keys on this sample."""
    states = [r["state"] for r in rows]
    keys = env.keys_batch(env.boards_from_states(states)).cpu().numpy()
    seen = {}
    for s, k in zip(states, map(tuple, keys)):
        if k in seen:
            assert seen[k] == s
        seen[k] = s
    assert len(set(seen.values())) == len(seen)
    by_state = {}
    for s, k in zip(states, map(tuple, keys)):
        assert by_state.setdefault(s, k) == k


def check_single_api(env):
    """The static_env-compatible scalar API on the reference's own smoke vectors."""
    init = env.INIT_STATE
    lm = env.get_legal_moves(init)
    assert lm == ('0001 0002 1022 1002 2042 2002 3041 4041 5041 6082 6042 7082 7062 8081 8082 1202 1222 1232 '
                  '1242 1252 1262 1211 1213 1214 1215 1216 1219 7222 7232 7242 7252 7262 7282 7271 7273 7274 '
                  '7275 7276 7279 0304 2324 4344 6364 8384').split()
    s1 = env.step(init, '0001')
    assert s1 == 'rkemsmek1/8r/1c5c1/p1p1p1p1p/9/9/P1P1P1P1P/1C5C1/9/RKEMSMEKR'
    assert env.step(s1, '1219') == 'rkemsmekr/9/1c7/p1p1p1p1p/9/9/P1P1P1P1P/1C5C1/R8/1KEMSMEcR'
    pl = env.state_to_planes(init)
    assert pl.shape == (14, 10, 9) and pl.sum() == 32
    assert pl.sum(axis=(1, 2)).tolist() == [5, 2, 2, 2, 2, 2, 1, 5, 2, 2, 2, 2, 2, 1]
    assert pl[6, 9, 4] == 1 and pl[13, 0, 4] == 1
    t = '4s4/9/4e4/p8/2e2R2p/P5E2/8P/9/9/4S1E2'
    assert env.done(t) == (False, 0, None)
    assert len(env.get_legal_moves(t)) == 24
    assert env.get_legal_moves('4s4/9/9/9/9/9/9/9/9/4S4') == ['4050', '4049', '4041', '4049', '4030', '4049']
    assert env.done('4s4/9/9/9/9/9/9/9/9/4S4') == (True, 1, None)
    assert env.done('9/9/9/9/9/9/9/9/9/4S4') == (True, 1, None)
    assert env.done('4s4/9/9/9/9/9/9/9/9/9', need_check=True) == (True, -1, None)
    assert env.new_step(init, '1219') == osenv.new_step(init, '1219')
    assert len(env.labels) == 2086 and env.labels == osenv.ActionLabelsRed
    try:
        env.step(init, '4445')
        assert False
    except ValueError:
        pass
    # empty batch is a no-op
    import torch
    e = torch.empty((0, 96), dtype=torch.uint8, device=env.device)
    mv, cnt = env.movegen_batch(e)
    assert mv.shape[0] == 0 and cnt.shape[0] == 0
