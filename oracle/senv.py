"""oracle/senv.py — CPU restatement of the reference rules engine.  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module;
the product path (chinesechess-alphazero_b200/) never does.

Restates cchess_alphazero/environment/static_env.py on a flat 90-int board (sq = y*9+x, y = 0 the
mover's back rank; 0 empty, 1..7 mover's P C R N E A K, 9..15 opponent's) and keeps the
reference's string API so tests read like calls into `senv`.  Pinned against the real reference
by tests/test_oracle_vs_reference.py (runs where /root/reference exists) and by the golden
vectors under tests/golden/ generated from the real reference (oracle/gen_golden.py).

Reference lines restated: INIT_STATE :9, done :14-77, step :79-86, new_step :88-98,
state_to_board :117-135, state_to_planes :137-156, board_to_state :196-213, fliped_state :245-254,
get_legal_moves :256-321, will_check_or_catch :390-421, get_catch_list :423-454,
be_catched :456-469, has_attack_chessman :471-479; letter maps light_env/common.py:32-64,
step vectors light_env/common.py:66-76; plane order lookup_tables.py:27-42.
"""
import numpy as np

INIT_STATE = 'rkemsmekr/9/1c5c1/p1p1p1p1p/9/9/P1P1P1P1P/1C5C1/9/RKEMSMEKR'

P, C, R, N, E, A, K = 1, 2, 3, 4, 5, 6, 7
OPP = 8
# state-string alphabet: UPPER = side to move; k = knight, e = elephant, m = advisor, s = king
_LETTER_TO_CODE = {'P': P, 'C': C, 'R': R, 'K': N, 'E': E, 'M': A, 'S': K}
_LETTER_TO_CODE.update({k.lower(): v | OPP for k, v in list(_LETTER_TO_CODE.items())})
_CODE_TO_LETTER = {v: k for k, v in _LETTER_TO_CODE.items()}

_STEPS = {
    K: ((0, -1), (1, 0), (0, 1), (-1, 0)),
    A: ((-1, -1), (1, -1), (-1, 1), (1, 1)),
    E: ((-2, -2), (2, -2), (2, 2), (-2, 2)),
    N: ((-1, -2), (1, -2), (2, -1), (2, 1), (1, 2), (-1, 2), (-2, 1), (-2, -1)),
    P: ((0, 1), (-1, 0), (1, 0)),
}


# ------------------------------------------------------------------ codec
def state_to_codes(state):
    """State string -> flat board.  String rows run from y = 9 down to y = 0."""
    b = [0] * 90
    y, x = 9, 0
    for ch in state:
        if ch == ' ':
            break
        if ch == '/':
            y -= 1
            x = 0
        elif ch.isdigit():
            x += int(ch)
        else:
            b[y * 9 + x] = _LETTER_TO_CODE[ch]
            x += 1
    return b


def codes_to_state(b):
    rows = []
    for y in range(9, -1, -1):
        s, gap = '', 0
        for x in range(9):
            c = b[y * 9 + x]
            if c == 0:
                gap += 1
            else:
                if gap:
                    s += str(gap)
                    gap = 0
                s += _CODE_TO_LETTER[c]
        if gap:
            s += str(gap)
        rows.append(s)
    return '/'.join(rows)


def flip_codes(b):
    return [(c ^ OPP) if c else 0 for c in reversed(b)]


def fliped_state(state):
    return codes_to_state(flip_codes(state_to_codes(state)))


def mv_str(f, t):
    return '%d%d%d%d' % (f % 9, f // 9, t % 9, t // 9)


def mv_sq(m):
    return int(m[1]) * 9 + int(m[0]), int(m[3]) * 9 + int(m[2])


# ------------------------------------------------------------------ move generation
def _own(c):
    return c != 0 and not (c & OPP)


def _ok(b, x, y):
    return 0 <= x <= 8 and 0 <= y <= 9 and not _own(b[y * 9 + x])


def _first_piece(b, x, y, dx, dy):
    """Coordinate of the first occupied square from (x,y) exclusive in direction (dx,dy);
    one step past the edge if none (x_board_from / y_board_from)."""
    x += dx
    y += dy
    while 0 <= x <= 8 and 0 <= y <= 9 and b[y * 9 + x] == 0:
        x += dx
        y += dy
    return x, y


def legal_moves_codes(b):
    """Ordered pseudo-legal (from, to) pairs of the side to move."""
    out = []
    for sq in range(90):
        c = b[sq]
        if not _own(c):
            continue
        x, y = sq % 9, sq // 9
        if c in (R, C):
            l = _first_piece(b, x, y, -1, 0)[0]
            r = _first_piece(b, x, y, 1, 0)[0]
            d = _first_piece(b, x, y, 0, -1)[1]
            u = _first_piece(b, x, y, 0, 1)[1]
            out += [(sq, y * 9 + xx) for xx in range(l + 1, x)]
            out += [(sq, y * 9 + xx) for xx in range(x + 1, r)]
            out += [(sq, yy * 9 + x) for yy in range(d + 1, y)]
            out += [(sq, yy * 9 + x) for yy in range(y + 1, u)]
            if c == C:   # one screen to jump
                l = _first_piece(b, l, y, -1, 0)[0] if l >= 0 else -2
                r = _first_piece(b, r, y, 1, 0)[0] if r <= 8 else 10
                d = _first_piece(b, x, d, 0, -1)[1] if d >= 0 else -2
                u = _first_piece(b, x, u, 0, 1)[1] if u <= 9 else 11
            for tx, ty in ((l, y), (r, y), (x, d), (x, u)):
                if _ok(b, tx, ty):
                    out.append((sq, ty * 9 + tx))
            continue
        fly = None
        if c == K:
            ux, uy = _first_piece(b, x, y, 0, 1)
            if uy <= 9 and b[uy * 9 + ux] == (K | OPP):
                fly = uy * 9 + ux
        for dx, dy in _STEPS[c]:
            tx, ty = x + dx, y + dy
            if not _ok(b, tx, ty):
                continue
            if c == P:
                if y < 5 and tx != x:
                    continue
            elif c in (N, E):
                if b[(y + int(dy / 2)) * 9 + x + int(dx / 2)] != 0:
                    continue
                if c == E and ty > 4:
                    continue
            else:
                if tx < 3 or tx > 5 or ty > 2:
                    continue
            out.append((sq, ty * 9 + tx))
            if fly is not None:
                out.append((sq, fly))
    return out


def get_legal_moves(state, board=None):
    return [mv_str(f, t) for f, t in legal_moves_codes(state_to_codes(state))]


# ------------------------------------------------------------------ transitions
def step_codes(b, f, t):
    nb = list(b)
    nb[t] = nb[f]
    nb[f] = 0
    return flip_codes(nb)


def step(state, action):
    b = state_to_codes(state)
    f, t = mv_sq(action)
    if b[f] == 0:
        raise ValueError(f"No chessman in {action}, state = {state}")
    return codes_to_state(step_codes(b, f, t))


def new_step(state, action):
    b = state_to_codes(state)
    f, t = mv_sq(action)
    if b[f] == 0:
        raise ValueError(f"No chessman in {action}, state = {state}")
    return codes_to_state(step_codes(b, f, t)), b[t] == 0


# ------------------------------------------------------------------ terminal test
def _last(b, code):
    pos = -1
    for sq in range(90):
        if b[sq] == code:
            pos = sq
    return pos


def done_codes(b, need_check=False):
    """(over, v, final_move or None, check) on a flat board."""
    opp_k = _last(b, K | OPP)
    own_k = _last(b, K)
    if opp_k < 0:
        return True, 1, None, False
    if own_k < 0:
        return True, -1, None, False
    over, v = False, 0
    if own_k == 0:          # the reference's [0, 0] "not found" sentinel
        over, v = True, -1
    elif opp_k == 0:
        over, v = True, 1
    elif own_k % 9 == opp_k % 9:
        x = own_k % 9
        if all(b[y * 9 + x] == 0 for y in range(own_k // 9 + 1, opp_k // 9)):
            over, v = True, 1
    final = None
    if not over:
        for f, t in legal_moves_codes(b):
            if t == opp_k:
                over, v, final = True, 1, (f, t)
                break
    check = False
    if not over and need_check:
        target = 89 - own_k
        check = any(t == target for _, t in legal_moves_codes(flip_codes(b)))
    return over, v, final, check


def done(state, turns=-1, need_check=False):
    if 's' not in state:
        return (True, 1, None)
    if 'S' not in state:
        return (True, -1, None)
    over, v, final, check = done_codes(state_to_codes(state), need_check)
    fm = mv_str(*final) if final else None
    return (over, v, fm, check) if need_check else (over, v, fm)


def has_attack_chessman(state):
    return any(ch.lower() in 'rkpc' for ch in state if ch.isalpha())


# ------------------------------------------------------------------ planes
def state_to_planes(state):
    planes = np.zeros((14, 10, 9), dtype=np.float32)
    b = state_to_codes(state)
    for sq in range(90):
        c = b[sq]
        if c:
            planes[(c - 2) if (c & OPP) else (c - 1), 9 - sq // 9, sq % 9] = 1
    return planes


def state_history_to_planes(state, history):
    """static_env.py:158-194: planes 0-13 = `state`, planes 14-27 = history[-5] (the position two plies
    earlier, same side to move) when the history list [.., state, move, state, move, state] holds >= 5 entries."""
    planes = np.zeros((28, 10, 9), dtype=np.float32)
    planes[:14] = state_to_planes(state)
    if history and len(history) >= 5:
        planes[14:] = state_to_planes(history[-5])
    return planes


# ------------------------------------------------------------------ UCI / FEN notation (static_env.py:224-228,380-388)
_FEN_TO_STATE = str.maketrans("nNbBaAkK", "kKeEmMsS")      # FEN knight/bishop/advisor/king -> state k/e/m/s (common.py:32-47)


def fen_to_state(fen):
    return fen.split(' ')[0].translate(_FEN_TO_STATE)


_STATE_TO_FEN = str.maketrans("kKeEmMsS", "nNbBaAkK")


def state_to_fen(state, turns):
    """static_env.py:215-243: FEN of the position; on black's turns (odd) the canonical state is turned back to the
    board's orientation (rows reversed, each row mirrored, colours swapped) and the side-to-move field becomes b."""
    fen = state.translate(_STATE_TO_FEN)
    if turns % 2 == 0:
        return f"{fen} w - - 0 {turns}"
    rows = ["".join(c.swapcase() for c in reversed(row)) for row in reversed(fen.split('/'))]
    return "/".join(rows) + f" b - - 0 {turns}"


def parse_ucci_move(move):
    return str(ord(move[0]) - ord('a')) + move[1] + str(ord(move[2]) - ord('a')) + move[3]


def to_uci_move(action):
    return chr(ord('a') + int(action[0])) + action[1] + chr(ord('a') + int(action[2])) + action[3]


# ------------------------------------------------------------------ repetition rules
def catch_set_codes(b, moves=None):
    res = set()
    if not moves:
        moves = legal_moves_codes(b)
    for f, t in moves:
        if b[t] == 0:
            continue
        nb = step_codes(b, f, t)
        back = 89 - t
        if any(tt == back for _, tt in legal_moves_codes(nb)):
            continue
        pf, pt = b[f], b[t]
        if pf == P and f // 9 <= 4:
            continue
        if pt == (P | OPP) and t // 9 > 4:
            continue
        if (pf & 7) == (pt & 7):
            continue
        res.add((pf, f, pt, t))
    return res


def will_check_or_catch_codes(b, f, t):
    after = step_codes(b, f, t)
    their_k = _last(after, K)
    again = flip_codes(after)
    moves = legal_moves_codes(again)
    target = 89 - their_k if their_k >= 0 else 89
    if any(tt == target for _, tt in moves):
        return True
    first = catch_set_codes(b)
    second = catch_set_codes(again, moves)
    return bool(second - first) and len(second) >= len(first)


def will_check_or_catch(ori_state, action):
    return will_check_or_catch_codes(state_to_codes(ori_state), *mv_sq(action))


def be_catched_codes(b, f):
    target = 89 - f
    return any(t == target for _, t in legal_moves_codes(flip_codes(b)))


def be_catched(state, mov):
    return be_catched_codes(state_to_codes(state), mv_sq(mov)[0])


# ------------------------------------------------------------------ labels (lookup_tables.py:50-134)
def flip_move(m):
    return '%d%d%d%d' % (8 - int(m[0]), 9 - int(m[1]), 8 - int(m[2]), 9 - int(m[3]))


def create_action_labels():
    labels = []
    for y0 in range(10):
        for x0 in range(9):
            dests = [(y0, t) for t in range(9)] + [(t, x0) for t in range(10)] + \
                    [(y0 + a, x0 + b) for a, b in ((-2, -1), (-1, -2), (-2, 1), (1, -2), (2, -1), (-1, 2), (2, 1), (1, 2))]
            for y1, x1 in dests:
                if (y0, x0) != (y1, x1) and 0 <= y1 < 10 and 0 <= x1 < 9:
                    labels.append('%d%d%d%d' % (x0, y0, x1, y1))
    adv_red = ['3041', '5041', '3241', '5241', '4130', '4150', '4132', '4152']
    ele_red = ['2002', '2042', '6042', '6082', '2402', '2442', '6442', '6482',
               '0220', '4220', '4260', '8260', '0224', '4224', '4264', '8264']
    labels += adv_red
    labels += ['3948', '5948', '3748', '5748', '4839', '4859', '4837', '4857']
    labels += ele_red
    labels += ['2907', '2947', '6947', '6987', '2507', '2547', '6547', '6587',
               '0729', '4729', '4769', '8769', '0725', '4725', '4765', '8765']
    return labels


ActionLabelsRed = create_action_labels()
