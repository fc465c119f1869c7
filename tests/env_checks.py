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


EXTREME_STATES = [
    '9/9/9/9/9/9/9/9/9/9',                                       # empty board: no kings at all
    '4s4/9/9/9/9/9/9/9/9/9',                                     # side to move has no king
    '9/9/9/9/9/9/9/9/9/4S4',                                     # opponent has no king
    '3s5/9/9/9/9/9/9/9/9/4S4',                                   # bare kings, different files
    'R2s4R/9/9/C7C/9/9/C7C/9/9/R3S3R',                           # rooks and cannons with long open lines (many moves)
    'rkemsmekr/9/1c5c1/p1p1p1p1p/9/9/P1P1P1P1P/1C5C1/9/RKEMSMEKR',
    '4s4/4P4/9/9/9/9/9/9/4p4/4S4',                               # pawns next to the kings
    '3ms4/4m4/4e4/9/2e6/6E2/9/4E4/4M4/4SM3',                     # only defenders left: no attacking piece
    'r1e1s1e1r/4m4/2k1m1k2/p1p1C1p1p/9/9/P1P1P1P1P/2K1C1K2/9/R1EMSME1R',   # cannon pinning through two screens
    '4s4/9/9/9/4R4/9/9/9/9/4S4',                                 # rook between the kings
    '5s3/9/9/9/9/9/9/9/5r3/3S5',
    'PPPPPPPPP/PPPPsPPPP/PPPPPPPPP/9/9/9/9/9/4S4/9',             # far more pawns than a real game can have
    'PPPPPPPPP/PPPPsPPPP/PPPPPPPPP/PPPPPPPPP/PPPP1PPPP/9/9/9/4S4/9',    # 44 own pieces: more than one warp of pieces
    'ppppppppp/pppp1pppp/ppppppppp/ppppppppp/pppp1pppp/R8/9/9/4S4/4s4',  # ... and of opposing ones
]


def random_boards(n, seed, own="RRKKEEMMCCPPPPP"):
    """Arbitrary (mostly unreachable) positions: both kings somewhere in their palaces, a random subset of the other 30
    pieces on random squares.  (Advisors / elephants on squares they can never reach make moves outside the 2086 action
    labels: fine for the rules, a KeyError for the reference's search - pass own="RRKKCCPPPPP" for search tests.)"""
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        sq = [None] * 90
        ks = [(y, x) for y in range(3) for x in range(3, 6)]
        y, x = ks[rng.randint(9)]
        sq[y * 9 + x] = 'S'
        y, x = ks[rng.randint(9)]
        sq[(9 - y) * 9 + x] = 's'
        for side in (own, own.lower()):
            for c in side:
                if rng.rand() < 0.6:
                    k = rng.randint(90)
                    if sq[k] is None:
                        sq[k] = c
        rows = []
        for y in range(9, -1, -1):
            row, gap = "", 0
            for x in range(9):
                c = sq[y * 9 + x]
                if c is None:
                    gap += 1
                else:
                    row += (str(gap) if gap else "") + c
                    gap = 0
            rows.append(row + (str(gap) if gap else ""))
        out.append("/".join(rows))
    return out


def check_random_boards(env, n=400, seed=12):
    states = random_boards(n, seed)
    boards = env.boards_from_states(states)
    mv, cnt = env.movegen_batch(boards)
    mv, cnt = mv.cpu().numpy().view(np.uint16), cnt.cpu().numpy()
    out, fm = env.done_batch(boards, need_check=True)
    out, fm = out.cpu().numpy(), fm.cpu().numpy().view(np.uint16)
    planes = env.planes_batch(boards).cpu().numpy()
    for i, s in enumerate(states):
        ref_moves = osenv.get_legal_moves(s)
        assert [u16_to_move(v) for v in mv[i, :cnt[i]]] == ref_moves, s
        d = osenv.done(s, need_check=True)
        assert (bool(out[i, 0]), int(out[i, 1])) == (d[0], d[1]) and (None if fm[i] == 0xFFFF else u16_to_move(fm[i])) == d[2], s
        if len(d) == 4:
            assert bool(out[i, 2]) == d[3], s
        assert (planes[i] == osenv.state_to_planes(s)).all(), s
        if i % 8 == 0 and ref_moves and not d[0]:
            m = ref_moves[(i // 8) % len(ref_moves)]
            assert env.new_step(s, m) == osenv.new_step(s, m), (s, m)
            assert env.will_check_or_catch(s, m) == osenv.will_check_or_catch(s, m), (s, m)
            assert env.be_catched(s, m) == osenv.be_catched(s, m), (s, m)


def check_extreme_positions(env):
    """Hand-made boundary positions: missing kings, empty board, maximum mobility, piece counts outside real games."""
    states = EXTREME_STATES
    boards = env.boards_from_states(states)
    mv, cnt = env.movegen_batch(boards)
    mv = mv.cpu().numpy().view(np.uint16)
    cnt = cnt.cpu().numpy()
    out, fm = env.done_batch(boards, need_check=True)
    out, fm = out.cpu().numpy(), fm.cpu().numpy().view(np.uint16)
    planes = env.planes_batch(boards).cpu().numpy()
    for i, s in enumerate(states):
        ref_moves = osenv.get_legal_moves(s)
        assert [u16_to_move(v) for v in mv[i, :cnt[i]]] == ref_moves, s
        d = osenv.done(s, need_check=True)
        assert (bool(out[i, 0]), int(out[i, 1])) == (d[0], d[1]), (s, out[i], d)
        assert (None if fm[i] == 0xFFFF else u16_to_move(fm[i])) == d[2], s
        if len(d) == 4:
            assert bool(out[i, 2]) == d[3], s
        assert (planes[i] == osenv.state_to_planes(s)).all(), s
        assert env.has_attack_chessman(s) == osenv.has_attack_chessman(s), s
        for m in ref_moves[:6]:
            assert env.new_step(s, m) == osenv.new_step(s, m), (s, m)
            if 's' in s and 'S' in s:
                assert env.will_check_or_catch(s, m) == osenv.will_check_or_catch(s, m), (s, m)
                assert env.be_catched(s, m) == osenv.be_catched(s, m), (s, m)
    assert cnt.max() >= 60                                        # the mobility case really is large
    # ragged / large batch: 20 000 boards in one call equal the same boards one by one
    import torch
    k = len(states)
    reps = 20000 // k + 1
    big = boards.repeat(reps, 1)[:20000]
    mv2, cnt2 = env.movegen_batch(big)
    assert torch.equal(cnt2.cpu(), torch.as_tensor(cnt).repeat(reps)[:20000])
    assert torch.equal(mv2[k:2 * k].cpu().view(torch.int16), torch.as_tensor(mv.view(np.int16)))
