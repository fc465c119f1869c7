"""oracle/selfplay.py — CPU restatement of the self-play game loop.  TEST INFRASTRUCTURE.

Restates worker/self_play.py:95-212 (SelfPlayWorker.start_game) for ONE game around an OraclePlayer, with the three
random decisions of the reference (resign lottery :102-105, move sampling player.py:195, store lottery :194-200)
injectable, so the device loop (csrc/cz_selfplay.cuh), which draws them from Philox streams, can be compared move
for move.  `DeviceDraws` re-implements those Philox draws and the device's sampling arithmetic in Python.

Pinned: with the random decisions taken from the reference's own generators (ref_worker_harness.ReferenceDraws) this
loop replays whole games of the REAL, unmodified SelfPlayWorker.start_game move for move (tests/golden/games_k1.json.gz,
tests/test_games_golden.py; live in tests/test_oracle_vs_reference.py).
"""
import math

import numpy as np

from . import senv
from .player import OraclePlayer


def play_game(pc, evaluate, draws, max_game_length=100, enable_resign_rate=0.5, env=senv, max_plies_guard=1000,
              use_history=False):
    """Returns dict(moves, value_red, turns, flags, store).  `draws` supplies resign_lottery(), choose(...), store_lottery()."""
    enable_resign = draws.resign_lottery() > enable_resign_rate
    player = OraclePlayer(pc, evaluate, env=env, enable_resign=enable_resign,
                          noise=(lambda n: 0.0) if (pc.noise_eps == 0 and not hasattr(draws, "choose_with_player")) else None,
                          use_history=use_history)     # the reference draws its Dirichlet sample even when eps == 0 (:304)
    # the game loop never hands its history to action() (self_play.py:124): history planes come from the search path only
    state = env.INIT_STATE
    history = [state]
    value, turns, game_over, final_move = 0, 0, False, None
    no_eat_count, check, no_act, increase_temp = 0, False, [], False
    flags = 0
    while not game_over and turns < max_plies_guard:
        player.search(state, no_act, increase_temp)
        player.increase_temp = increase_temp
        policy, resign = player.calc_policy(state, turns, no_act)
        if resign:
            value = -1
            flags |= 1
            break
        node = player.tree[state]
        if hasattr(draws, "choose_with_player"):            # the reference's own np.random.choice (oracle/ref_worker_harness.py)
            action = draws.choose_with_player(player, state, turns, no_act, increase_temp)
        else:
            action = draws.choose(node, no_act, turns, increase_temp, pc)
        history.append(action)
        state, no_eat = env.new_step(state, action)
        turns += 1
        no_eat_count = no_eat_count + 1 if no_eat else 0
        history.append(state)
        if no_eat_count >= 120 or turns / 2 >= max_game_length:
            game_over, value = True, 0
            flags |= 2
        else:
            game_over, value, final_move, check = env.done(state, need_check=True)
            if not game_over:
                if not env.has_attack_chessman(state):
                    game_over, value = True, 0
                    flags |= 2
            increase_temp = False
            no_act = []
            if not game_over and not check and state in history[:-1]:
                free_move = 0
                for i in range(len(history) - 1):
                    if history[i] == state:
                        if env.will_check_or_catch(state, history[i + 1]):
                            no_act.append(history[i + 1])
                        elif not env.be_catched(state, history[i + 1]):
                            increase_temp = True
                            free_move += 1
                            if free_move >= 3:
                                game_over, value = True, 0
                                flags |= 2
                                break
    if final_move:
        history.append(final_move)
        state = env.step(state, final_move)
        turns += 1
        value = -value
        history.append(state)
    if turns % 2 == 1:
        value = -value
    store = True
    if turns < 10:
        store = draws.store_lottery() > 0.9
    return {"moves": [history[2 * i + 1] for i in range(turns)], "value_red": value, "turns": turns, "flags": flags,
            "store": store, "final_state": state}


# ------------------------------------------------------------------ the device's Philox draws, restated
def philox4x32(c, k):
    c0, c1, c2, c3 = c
    k0, k1 = k
    M = 0xFFFFFFFF
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        n0 = ((p1 >> 32) ^ c1 ^ k0) & M
        n1 = p1 & M
        n2 = ((p0 >> 32) ^ c3 ^ k1) & M
        n3 = p0 & M
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + 0x9E3779B9) & M
        k1 = (k1 + 0xBB67AE85) & M
    return [c0, c1, c2, c3]


def philox_first_word(seed, rank, game, purpose, index):
    """First Rng::next() of the stream (csrc/cz_tree.cuh Rng): buf[3] of the first Philox block."""
    M = 0xFFFFFFFF
    k0 = ((seed & M) ^ ((rank * 0x632BE5AB) & M)) & M
    k1 = (((seed >> 32) & M) + 0x1234567 * rank) & M
    return philox4x32((game & M, purpose & M, index & M, 0), (k0, k1))[3]


def philox_uniform(seed, rank, game, purpose, index):
    """First Rng::uniform() of the stream (csrc/cz_tree.cuh Rng): buf[3] is the high word, buf[2] the low word."""
    M = 0xFFFFFFFF
    k0 = ((seed & M) ^ ((rank * 0x632BE5AB) & M)) & M
    k1 = (((seed >> 32) & M) + 0x1234567 * rank) & M
    out = philox4x32((game & M, purpose & M, index & M, 0), (k0, k1))
    a, b = out[3], out[2]
    return (float(((a << 32) | b) >> 11) + 0.5) * (1.0 / 9007199254740992.0)


class DeviceDraws:
    """The random decisions of one game slot exactly as csrc/cz_selfplay.cuh takes them."""

    def __init__(self, seed, rank, game, games_started, label_of):
        self.seed, self.rank, self.game, self.started = seed, rank, game, games_started
        self.label_of = label_of            # move string -> label index

    def resign_lottery(self):
        return philox_uniform(self.seed, self.rank, self.game, 3, self.started)

    def store_lottery(self):
        return philox_uniform(self.seed, self.rank, self.game, 4, self.started)

    def playouts(self, lo, hi):
        """Arena: simulations per move of the game this slot starts (selfplay_start_game; `game` must be the game's first
        slot, i.e. slot % M, so that both player slots agree)."""
        return (lo + philox_first_word(self.seed, self.rank, self.game, 5, self.started) % (hi - lo + 1)) * 100

    def choose(self, node, no_act, turns, increase_temp, pc):
        moves = node.legal_moves
        L = len(moves)
        ban = [bool(no_act) and m in no_act for m in moves]
        n = [0 if ban[i] else (node.a[moves[i]].n if moves[i] in node.a else 0) for i in range(L)]
        lab = [self.label_of[m] for m in moves]
        sum_n = float(sum(n))
        tau = 0.0
        if turns < 30 and pc.tau_decay_rate != 0.0:
            tau = math.pow(pc.tau_decay_rate, float(turns + 1))
        if tau < 0.1:
            tau = 0.0
        if increase_temp:
            tau = 0.5
        if tau == 0.0:
            best = None
            for i in range(L):
                if ban[i]:
                    continue
                key = (-float(n[i]), lab[i])
                if best is None or key < best[0]:
                    best = (key, i)
            return moves[best[1]]
        w = [math.pow(n[i] / sum_n, 1.0 / tau) if n[i] > 0 else 0.0 for i in range(L)]
        lane = [0.0] * 32                                   # per-lane partial sums, then the xor butterfly
        for l in range(32):
            s = 0.0
            for c in range(4):
                i = c * 32 + l
                s = (s + (w[i] if i < L else 0.0)) if c else (w[i] if i < L else 0.0)
            lane[l] = s
        for m in (16, 8, 4, 2, 1):
            lane = [lane[l] + lane[l ^ m] for l in range(32)]
        tot = lane[0]
        u = philox_uniform(self.seed, self.rank, self.game, 2, self.started * 1024 + turns) * tot
        best_c, bi, fb_l, fb = 1e300, -1, -1, -1
        for i in range(L):
            if w[i] <= 0.0:
                continue
            before = 0.0
            for j in range(L):
                if lab[j] < lab[i]:
                    before += w[j]
            if before + w[i] > u and before < best_c:
                best_c, bi = before, i
            if lab[i] > fb_l:
                fb_l, fb = lab[i], i
        return moves[bi if bi >= 0 else fb]
