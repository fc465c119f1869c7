"""`CChessPlayer` drop-in (reference: cchess_alphazero/agent/player.py:35-470) backed by the GPU engine.

Same constructor and `action()` signature and return values as the reference class.  The search
(MCTS_search / select / expand / update_tree, player.py:198-373) runs in the CUDA kernels behind
`cz_search_*`; this class only does what the reference does once per move on the host: the visit-count
policy, temperature and the `np.random.choice` draw (player.py:187-196,375-406,453-470), with numpy, so
that the global `np.random` stream is consumed exactly like the reference consumes it:

  * root Dirichlet noise: the reference draws `np.random.dirichlet(alpha*ones(L))[0]` once per legal
    move per root selection (player.py:304).  The player pre-draws that sequence from the current
    `np.random` state into a device table, lets the engine consume a prefix, then rewinds the generator
    and re-draws exactly the consumed prefix, leaving the stream where the reference would leave it.
  * evaluation: if `pipes` is given the leaves are sent through it with the reference wire protocol
    (list of float32[14,10,9] ([28,10,9] with use_history) -> list of (float32[2086], float), api.py:48-74), so the player works
    against an unmodified CChessModelAPI; otherwise the engine's built-in tensor-core network is used.
"""
import numpy as np
import torch

from .engine import Engine
from .env import StaticEnv
from .lib import get_lib


class CChessPlayer:
    def __init__(self, config, search_tree=None, pipes=None, play_config=None, enable_resign=False, debugging=False,
                 uci=False, use_history=False, side=0, lib=None, device=None, weights=None, exact_noise=True):
        self.use_history = use_history          # 28 input planes (static_env.py:158-194, player.py:326-334)
        self.config = config
        self.play_config = play_config or config.play
        self.lib = lib or get_lib()
        self.env = StaticEnv(self.lib, device)
        self.labels = self.env.labels
        self.labels_n = len(self.labels)
        self.move_lookup = {m: i for i, m in enumerate(self.labels)}
        self.pipe = pipes
        self.enable_resign = enable_resign
        self.debugging = debugging
        self.uci = uci
        self.side = side
        self.increase_temp = False
        self.no_act = None
        self.root_state = None
        self.tree = {}                 # the tree lives in device memory; kept for attribute compatibility
        self.debug = {}
        self.search_results = {}
        self.done_tasks = 0
        self.exact_noise = exact_noise
        pc = self.play_config
        mc = getattr(config, "model", None)
        use_nn = pipes is None
        self.engine = Engine(
            self.lib, self.env.device, n_games=1, sims_per_move=pc.simulation_num_per_move,
            leaves_per_round=config.play.search_threads, virtual_loss=config.play.virtual_loss,
            noise_mode=0 if exact_noise else 1, c_puct=pc.c_puct, noise_eps=pc.noise_eps,
            dirichlet_alpha=pc.dirichlet_alpha, tau_decay_rate=pc.tau_decay_rate,
            resign_threshold=getattr(pc, "resign_threshold", -1.0), min_resign_turn=getattr(pc, "min_resign_turn", 0),
            max_game_length=getattr(pc, "max_game_length", 100),
            max_nodes_per_game=max(4096, 8 * pc.simulation_num_per_move),
            nn_filters=mc.cnn_filter_num if (use_nn and mc) else 0, nn_blocks=mc.res_layer_num if (use_nn and mc) else 0,
            nn_value_fc=mc.value_fc_size if (use_nn and mc) else 256, use_history=use_history)
        if use_nn:
            if weights is None:
                raise ValueError("CChessPlayer without pipes needs `weights` (Keras-named tensors) for the built-in network")
            self.engine.set_weights(weights)
        self._fresh = True

    # ---- reference API
    def close(self, wait=True):
        if self.engine is not None:
            self.engine.close()
            self.engine = None

    def action(self, state, turns, no_act=None, depth=None, infinite=False, hist=None, increase_temp=False):
        if infinite:
            raise NotImplementedError("infinite analysis (uci.py) is outside the built hot path")
        pc = self.play_config
        eng = self.engine
        self.root_state = state
        self.no_act = no_act
        self.increase_temp = increase_temp
        if self._fresh:
            eng.reset([state])
            self._fresh = False
        else:
            eng.set_root(0, state)
        noise, rng_state, n_moves = None, None, 0
        if self.exact_noise and pc.noise_eps != 0:
            n_moves = len(self.env.get_legal_moves(state))
            sims = depth if depth else pc.simulation_num_per_move
            rng_state = np.random.get_state()
            cap = (sims + 2 * self.engine.K + 2) * max(n_moves, 1)
            alpha = pc.dirichlet_alpha * np.ones(max(n_moves, 1))
            noise = np.array([np.random.dirichlet(alpha)[0] for _ in range(cap)], dtype=np.float64)[None, :]
        opts = eng.make_opts(no_act=[list(no_act)] if no_act else None, increase_temp=[1 if increase_temp else 0],
                             noise=noise, sims_override=int(depth) if depth else 0,
                             hist=[list(hist)] if (self.use_history and hist) else None)
        if self.pipe is not None:
            eng.search_external(self._evaluate_through_pipe, opts)
        else:
            eng.search(opts)
        root = eng.root(0)
        self.done_tasks += root["sims_run"]
        if rng_state is not None:
            np.random.set_state(rng_state)
            alpha = pc.dirichlet_alpha * np.ones(max(n_moves, 1))
            for _ in range(root["noise_used"]):
                np.random.dirichlet(alpha)
        policy, resign = self.calc_policy(root, turns, no_act)
        if resign:
            return None, list(policy)
        if no_act is not None:
            for act in no_act:
                policy[self.move_lookup[act]] = 0
        my_action = int(np.random.choice(range(self.labels_n), p=self.apply_temperature(policy, turns)))
        return self.labels[my_action], list(policy)

    def close_and_return_action(self, state, turns, no_act=None):
        """player.py:88-106 (used by the UCI front end to stop an ongoing search): answer from the tree as it stands."""
        root = self.engine.root(0)
        policy, resign = self.calc_policy(root, turns, no_act)
        if resign:
            return None
        if no_act is not None:
            for act in no_act:
                policy[self.move_lookup[act]] = 0
        my_action = int(np.random.choice(range(self.labels_n), p=self.apply_temperature(policy, turns)))
        value = self.debug.get(state, (None, 0))[1]
        return self.labels[my_action], value, self.done_tasks // 100

    # ---- host-side tail of action(): player.py:375-406
    def calc_policy(self, root, turns, no_act):
        policy = np.zeros(self.labels_n)
        max_q_value = -100
        for mov, n, w in zip(root["moves"], root["n"], root["w"]):
            policy[self.move_lookup[mov]] = n
            if no_act and mov in no_act:
                policy[self.move_lookup[mov]] = 0
                continue
            q = w / n if n != 0 else 0
            if q > max_q_value:
                max_q_value = q
        pc = self.play_config
        if max_q_value < getattr(pc, "resign_threshold", -1e9) and self.enable_resign and turns > getattr(pc, "min_resign_turn", 0):
            return policy, True
        if self.debugging:                               # player.py:397-403: the five most visited moves
            order = sorted(range(len(root["moves"])), key=lambda i: root["n"][i], reverse=True)[:5]
            for i in order:
                n, w = root["n"][i], root["w"][i]
                if not (no_act and root["moves"][i] in no_act):
                    self.search_results[root["moves"][i]] = (n, w / n if n else 0, root["p"][i])
        policy /= np.sum(policy)
        return policy, False

    # ---- player.py:453-470
    def apply_temperature(self, policy, turn):
        pc = self.play_config
        evaluate = bool(getattr(getattr(self.config, "opts", None), "evaluate", False))
        if turn < 30 and pc.tau_decay_rate != 0:
            tau = np.power(pc.tau_decay_rate, turn + 1)
        else:
            tau = 0
        if tau < 0.1 or (turn >= 4 and evaluate):
            tau = 0
        if self.increase_temp and not evaluate:
            tau = 0.5
        if tau == 0:
            ret = np.zeros(self.labels_n)
            ret[np.argmax(policy)] = 1.0
            return ret
        ret = np.power(policy, 1 / tau)
        ret /= np.sum(ret)
        return ret

    # ---- the reference wire protocol (player.py:118-120,131-140 <-> api.py:48-74)
    def _evaluate_through_pipe(self, planes):
        out_p, out_v = [], []
        for i in range(0, len(planes), 256):
            chunk = planes[i:i + 256]
            self.pipe.send([np.ascontiguousarray(p, dtype=np.float32) for p in chunk])
            rets = self.pipe.recv()
            for p, v in rets:
                out_p.append(np.asarray(p, dtype=np.float32))
                out_v.append(np.float32(v))
        return np.stack(out_p), np.asarray(out_v, dtype=np.float32)
