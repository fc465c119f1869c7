"""oracle/player.py — CPU restatement of the reference MCTS player.  TEST INFRASTRUCTURE.

Restates cchess_alphazero/agent/player.py (CChessPlayer) as a deterministic, single-threaded
program: the reference's thread pool is replaced by ONE worker draining a FIFO task queue and the
network reply is delivered whenever that queue runs dry ("canonical
schedule").  With search_threads = 1 this is exactly what the real player does, which is how the
restatement is pinned (tests/test_oracle_vs_reference.py, tests/golden/mcts_k1.json.gz).

Reference lines restated: VisitState/ActionState :17-33, action :145-196, MCTS_search :198-260,
select_action_q_and_u :262-320, expand_and_evaluate :322-338, update_tree :340-373,
calc_policy :375-406, apply_temperature :453-470.

Arithmetic is written with the same numpy / Python scalar types as the reference so that numpy 2
promotion rules (the oracle's pinned semantics) produce identical bits:
priors are np.float32, W/Q Python floats, U mixes float32 and float64 exactly as :287-306 do.
"""
from collections import deque

import numpy as np

from . import senv


class Edge:
    __slots__ = ("n", "w", "q", "p")

    def __init__(self):
        self.n = 0
        self.w = 0
        self.q = 0
        self.p = 0


class Node:
    __slots__ = ("a", "sum_n", "visit", "p", "legal_moves", "waiting")

    def __init__(self):
        self.a = {}            # move -> Edge, in insertion order like the reference's defaultdict
        self.sum_n = 0
        self.visit = []        # parked simulations (their histories)
        self.p = None
        self.legal_moves = None
        self.waiting = False

    def edge(self, mov):
        e = self.a.get(mov)
        if e is None:
            e = self.a[mov] = Edge()
        return e


class PlayConfig:
    """The fields of config.play / play_config the player reads (configs/*.py PlayConfig)."""

    def __init__(self, simulation_num_per_move=800, search_threads=10, c_puct=1.5, noise_eps=0.15,
                 dirichlet_alpha=0.2, tau_decay_rate=0.9, virtual_loss=3, resign_threshold=-0.98,
                 min_resign_turn=40):
        self.simulation_num_per_move = simulation_num_per_move
        self.search_threads = search_threads
        self.c_puct = c_puct
        self.noise_eps = noise_eps
        self.dirichlet_alpha = dirichlet_alpha
        self.tau_decay_rate = tau_decay_rate
        self.virtual_loss = virtual_loss
        self.resign_threshold = resign_threshold
        self.min_resign_turn = min_resign_turn


class OraclePlayer:
    """evaluate(list_of_states) -> list of (policy float32[2086], float value), same order.
    use_history (player.py:45,326-334): evaluate(list_of_states, list_of_history_states_or_None) instead; the second
    list holds, per leaf, the state whose planes fill input planes 14-27 (None = zero planes)."""

    def __init__(self, play_config, evaluate, env=senv, tree=None, enable_resign=False, noise=None,
                 evaluate_mode=False, use_history=False, uci=False, side=0, debugging=False):
        self.pc = play_config
        self.evaluate = evaluate
        self.env = env
        self.tree = tree if tree is not None else {}
        self.labels = env.ActionLabelsRed
        self.move_lookup = {m: i for i, m in enumerate(self.labels)}
        self.enable_resign = enable_resign
        self.use_history = use_history
        self.uci, self.side, self.debugging = uci, side, debugging
        self.debug = {}                         # state -> (p, v) of every evaluated position when debugging (:349-350)
        self.info = []                          # (depth, score, pv string) of every `info depth` line (:408-450)
        self.done_tasks = 0
        self.evaluate_mode = evaluate_mode      # config.opts.evaluate
        # noise(move_count) -> one Dirichlet(alpha * 1_n)[0] draw; default = the reference's call
        self.noise = noise or (lambda n: np.random.dirichlet(self.pc.dirichlet_alpha * np.ones(n))[0])
        self.root_state = None
        self.no_act = None
        self.increase_temp = False
        self.queue = deque()
        self.buffer = []                        # (state, history) awaiting evaluation
        self.num_task = 0
        self.stats = {"sims": 0, "positions": 0, "batches": 0, "noise_draws": 0, "path_edges": 0, "no_network": 0}

    # ---- action(): player.py:145-196
    def search(self, state, no_act=None, increase_temp=False, depth=None, hist=None, turns=0, infinite=False, stop=None):
        """stop(): polled between rounds when `infinite` (the reference's close_and_return_action sets job_done)."""
        self.root_state = state
        self.no_act = no_act
        self.increase_temp = increase_temp
        if hist and len(hist) >= 5:                          # :150-151
            hist = hist[-5:]
        done = self.tree[state].sum_n if state in self.tree else 0
        if no_act or increase_temp or done == self.pc.simulation_num_per_move:
            done = 0
        self.done_tasks = done
        num_task = self.pc.simulation_num_per_move - done
        if depth:
            num_task = depth - done if depth > done else 0
        if infinite:
            num_task = 100000
        shown = 0
        if num_task > 0:
            k = self.pc.search_threads
            all_tasks = num_task
            batch = all_tasks // k + (1 if all_tasks % k else 0)
            for it in range(batch):
                if stop is not None and stop():
                    break
                self.num_task = min(k, all_tasks - k * it)
                self.done_tasks += self.num_task
                self.stats["sims"] += self.num_task
                for _ in range(self.num_task):
                    self.queue.append(("search", state, [state], hist))
                self._drain_until_round_done()
                if self.uci and shown != self.done_tasks // 100:             # :180-184
                    shown = self.done_tasks // 100
                    self.info.append(self.depth_info(state, turns, self.debug[state][1], no_act))

    # ---- print_depth_info: player.py:408-450 (returns (depth, score, " m1 m2 ...") instead of printing; time and nps
    #      are wall-clock and left out)
    def depth_info(self, state, turns, value, no_act):
        env = self.env
        pv = ""
        i = 0
        root = True
        while i < 20:
            node = self.tree.get(state)
            if node is None or len(node.a) == 0:
                break
            bestmove, n = None, 0
            for mov, a in node.a.items():
                if a.n >= n:
                    if root and no_act and mov in no_act:
                        continue
                    n, bestmove = a.n, mov
            if bestmove is None:
                break
            state = env.step(state, bestmove)
            root = False
            if turns % 2 == 1:
                bestmove = env.flip_move(bestmove)
            pv += " " + env.to_uci_move(bestmove)
            i += 1
            turns += 1
        if state in self.debug:
            _, value = self.debug[state]
            if turns % 2 != self.side:
                value = -value
        return (self.done_tasks // 100, int(value * 1000), pv)

    def action(self, state, turns, no_act=None, depth=None, increase_temp=False, hist=None, infinite=False, stop=None):
        self.search(state, no_act, increase_temp, depth, hist, turns, infinite, stop)
        policy, resign = self.calc_policy(state, turns, no_act)
        if resign:
            return None, list(policy)
        if no_act is not None:
            for act in no_act:
                policy[self.move_lookup[act]] = 0
        my_action = int(np.random.choice(range(len(self.labels)), p=self.apply_temperature(policy, turns)))
        return self.labels[my_action], list(policy)

    # ---- the single worker + the prediction round trip (sender/receiver, :108-143)
    def _drain_until_round_done(self):
        while self.num_task > 0:
            while self.queue:
                t = self.queue.popleft()
                if t[0] == "search":
                    self._mcts_search(t[1], t[2], t[3])
                else:
                    self._update_tree(t[1], t[2], t[3])
            if self.num_task <= 0:
                break
            if not self.buffer:
                raise RuntimeError("oracle player: deadlock (no queued work, no pending evaluation)")
            batch = self.buffer[:256]
            if self.use_history:
                rets = self.evaluate([s for s, _, _ in batch], [h for _, _, h in batch])
            else:
                rets = self.evaluate([s for s, _, _ in batch])
            self.stats["positions"] += len(batch)
            self.stats["batches"] += 1
            for (s, hist, _), (p, v) in zip(batch, rets):
                self.queue.append(("update", p, float(v), hist))
            self.buffer = self.buffer[len(batch):]

    # ---- MCTS_search: player.py:198-260
    def _mcts_search(self, state, history, real_hist=None):
        # real_hist: the `hist` argument of action(); only descents that START at the root carry it, and they use it
        # for every leaf they expand whatever its depth (is_root_node is never cleared inside the loop, :198-221);
        # simulations resumed by update_tree (:351-352) come without it.
        env = self.env
        while True:
            game_over, v, _ = env.done(state)
            if game_over:
                self.queue.append(("update", None, v * 2, history))
                return
            if state not in self.tree:
                node = self.tree[state] = Node()
                node.sum_n = 1
                node.legal_moves = env.get_legal_moves(state)
                node.waiting = True
                src = real_hist if real_hist else history  # expand_and_evaluate :322-338
                hist_state = src[-5] if (self.use_history and len(src) >= 5) else None
                self.buffer.append((state, history, hist_state))
                return
            if state in history[:-1]:
                for i in range(len(history) - 1):
                    if history[i] == state:
                        if env.will_check_or_catch(state, history[i + 1]):
                            self.queue.append(("update", None, -1, history))
                        elif env.be_catched(state, history[i + 1]):
                            self.queue.append(("update", None, 1, history))
                        else:
                            self.queue.append(("update", None, 0, history))
                        break
                return
            node = self.tree[state]
            if node.waiting:
                node.visit.append(history)
                return
            sel = self._select(state)
            vl = self.pc.virtual_loss
            node.sum_n += 1
            e = node.edge(sel)
            e.n += vl
            e.w -= vl
            e.q = e.w / e.n
            history.append(sel)
            state = env.step(state, sel)
            history.append(state)

    # ---- select_action_q_and_u: player.py:262-320
    def _select(self, state):
        is_root = self.root_state == state
        node = self.tree[state]
        legal = node.legal_moves
        if node.p is not None:
            all_p = 0
            for mov in legal:
                mov_p = node.p[self.move_lookup[mov]]
                node.edge(mov).p = mov_p
                all_p += mov_p
            if all_p == 0:
                all_p = 1
            for mov in legal:
                node.a[mov].p /= all_p
            node.p = None
        xx_ = np.sqrt(node.sum_n + 1)
        e_ = self.pc.noise_eps
        c_puct = self.pc.c_puct
        best_score, best = -99999999, None
        n_moves = len(legal)
        for mov in legal:
            if is_root and self.no_act and mov in self.no_act:
                continue
            a = node.edge(mov)
            p_ = a.p
            if is_root:
                self.stats["noise_draws"] += 1
                p_ = (1 - e_) * p_ + e_ * self.noise(n_moves)
            score = a.q + c_puct * p_ * xx_ / (1 + a.n)
            if a.q > (1 - 1e-7):
                best = mov
                break
            if score >= best_score:
                best_score, best = score, mov
        return best

    # ---- update_tree: player.py:340-373
    def _update_tree(self, p, v, history):
        self.stats["path_edges"] += len(history) // 2
        self.stats["no_network"] += 1 if p is None else 0
        state = history.pop()
        if p is not None:
            node = self.tree[state]
            node.p = p
            node.waiting = False
            if self.debugging:
                self.debug[state] = (p, v)
            for hist in node.visit:
                self.queue.append(("search", state, hist, None))
            node.visit = []
        vl = self.pc.virtual_loss
        while len(history) > 0:
            action = history.pop()
            state = history.pop()
            v = -v
            a = self.tree[state].edge(action)
            a.n += 1 - vl
            a.w += v + vl
            a.q = a.w * 1.0 / a.n
        self.num_task -= 1

    # ---- calc_policy: player.py:375-406
    def calc_policy(self, state, turns, no_act):
        node = self.tree[state]
        policy = np.zeros(len(self.labels))
        max_q = -100
        for mov, a in node.a.items():
            policy[self.move_lookup[mov]] = a.n
            if no_act and mov in no_act:
                policy[self.move_lookup[mov]] = 0
                continue
            if a.q > max_q:
                max_q = a.q
        if max_q < self.pc.resign_threshold and self.enable_resign and turns > self.pc.min_resign_turn:
            return policy, True
        policy /= np.sum(policy)
        return policy, False

    # ---- apply_temperature: player.py:453-470
    def apply_temperature(self, policy, turn):
        if turn < 30 and self.pc.tau_decay_rate != 0:
            tau = np.power(self.pc.tau_decay_rate, turn + 1)
        else:
            tau = 0
        if tau < 0.1 or (turn >= 4 and self.evaluate_mode):
            tau = 0
        if self.increase_temp and not self.evaluate_mode:
            tau = 0.5
        if tau == 0:
            ret = np.zeros(len(self.labels))
            ret[np.argmax(policy)] = 1.0
            return ret
        ret = np.power(policy, 1 / tau)
        ret /= np.sum(ret)
        return ret


# ------------------------------------------------------------------ deterministic stand-in network
def fake_eval_from_planes(planes):
    """Deterministic pseudo-network used by parity tests on BOTH sides (reference player, oracle, GPU
    engine): policy/value are a pure function of the 14x10x9 one-hot planes.  Integer hashing then an
    exact int->float32 conversion, so every implementation gets identical bits."""
    idx = np.flatnonzero(np.asarray(planes).reshape(-1)).astype(np.uint64)
    h = np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        for i in idx:
            h = (h ^ (i + np.uint64(0x7F4A7C15))) * np.uint64(0xBF58476D1CE4E5B9)
            h ^= h >> np.uint64(29)
        k = np.arange(2086, dtype=np.uint64)
        z = (h + k * np.uint64(0x94D049BB133111EB))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    raw = ((z >> np.uint64(40)) & np.uint64(0xFFFF)).astype(np.float32) + np.float32(1.0)   # 1..65536
    raw = raw * raw                                    # a little peakier than uniform
    policy = (raw / raw.sum(dtype=np.float32)).astype(np.float32)
    v = np.float32((int(h >> np.uint64(11)) % 2001 - 1000) / 1000.0) * np.float32(0.9)
    return policy, float(v)


def fake_evaluate_states(states, env=senv):
    return [fake_eval_from_planes(env.state_to_planes(s)) for s in states]


def fake_evaluate_states_hist(states, hist_states, env=senv):
    return [fake_eval_from_planes(env.state_history_to_planes(s, [h, None, None, None, s] if h else None))
            for s, h in zip(states, hist_states)]
