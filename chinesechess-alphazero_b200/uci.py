"""UCI / UCCI front end (reference: cchess_alphazero/uci.py:40-331) on top of the drop-in `CChessPlayer`.

Same command set and output as the reference: `uci`, `ucinewgame`, `setoption name gpu|Threads value x`, `isready`,
`position {fen <fen> | startpos} [moves ...]`, `fen ...`, `go [depth x] [movetime|time x] [wtime x] [btime x] [infinite]`,
`stop`, `quit`; answers `info depth .. score .. time .. pv .. nps ..` while thinking (printed by the player,
player.py:408-450), then `info depth .. score .. time .. nps ..` and `bestmove <m> [ponder <m>]` (uci.py:293-327).

Differences that do not show on the wire: one search engine on the GPU instead of a thread pool; by default the leaves
are evaluated by the built-in tensor-core network (`use_pipes=True` goes through CChessModelAPI's pipe like uci.py:205).
"""
import sys
from threading import Thread, Timer
from time import time

from .env import INIT_STATE, StaticEnv, fen_to_state, flip_move, parse_ucci_move, to_uci_move
from .lib import get_lib
from .player import CChessPlayer


class UCI:
    def __init__(self, config, model=None, lib=None, device=None, stdin=None, stdout=None, use_pipes=False, pipes_factory=None,
                 infinite_capacity=200000):
        self.config = config
        self.lib = lib or get_lib()
        self.device = device
        self.env = StaticEnv(self.lib, device)
        self.stdin = stdin or sys.stdin
        self.stdout = stdout or sys.stdout
        self.model = model
        self.use_pipes = use_pipes or pipes_factory is not None
        self.pipes_factory = pipes_factory       # () -> Connection; default: model.get_pipes(need_reload=False)
        self.infinite_capacity = infinite_capacity
        self.args = None
        self.state = None
        self.is_red_turn = None
        self.player = None
        self.is_ready = False
        self.remain_time = None
        self.history = None
        self.turns = 0
        self.start_time = None
        self.t = None
        self.use_history = False
        self.search_worker = None

    def _print(self, text):
        print(text, file=self.stdout)
        self.stdout.flush()

    # ---- uci.py:59-69
    def main(self):
        for line in self.stdin:
            cmd = line.strip()
            if not cmd:
                continue
            cmds = cmd.split(' ')
            self.args = cmds[1:]
            method = getattr(self, 'cmd_' + cmds[0], None)
            if method is not None:
                if method() == "quit":
                    return

    # ---- uci.py:71-88
    def cmd_uci(self):
        self._print('id name CCZero')
        self._print('id author https://cczero.org')
        self._print('id version 2.4')
        self._print('option name gpu spin default 0 min 0 max 7')
        self._print('option name Threads spin default 10 min 0 max 1024')
        self._print('uciok')
        self.use_history = self.load_model()
        self.is_ready = True
        self.turns = 0
        self.remain_time = None
        self.state = INIT_STATE
        self.history = [self.state]
        self.is_red_turn = True

    def cmd_ucinewgame(self):
        self.state = INIT_STATE
        self.history = [self.state]
        self.is_ready = True
        self.is_red_turn = True

    # ---- uci.py:97-110
    def cmd_setoption(self):
        if len(self.args) > 3:
            name = self.args[1]
            if name == 'gpu':
                self.device = f"cuda:{int(self.args[3])}" if self.lib.is_cuda else self.device
            if name == 'Threads':
                self.config.play.search_threads = int(self.args[3])

    def cmd_isready(self):
        if self.is_ready:
            self._print('readyok')

    # ---- uci.py:118-169
    def cmd_position(self):
        if not self.is_ready:
            return
        move_idx = -1
        if len(self.args) > 0:
            if self.args[0] == 'fen':
                try:
                    self.state = fen_to_state(self.args[1])
                    self.env.get_legal_moves(self.state)
                except Exception:
                    return
                self.history = [self.state]
                if self.args[2] == 'b':
                    self.state = self.env.fliped_state(self.state)
                    self.is_red_turn = False
                    self.turns = (int(self.args[6]) - 1) * 2 + 1
                else:
                    self.is_red_turn = True
                    self.turns = (int(self.args[6]) - 1) * 2
                if len(self.args) > 7 and self.args[7] == 'moves':
                    move_idx = 8
            elif self.args[0] == 'startpos':
                self.state = INIT_STATE
                self.is_red_turn = True
                self.history = [self.state]
                self.turns = 0
                if len(self.args) > 1 and self.args[1] == 'moves':
                    move_idx = 2
            elif self.args[0] == 'moves':
                move_idx = 1
        else:
            self.state = INIT_STATE
            self.is_red_turn = True
            self.history = [self.state]
            self.turns = 0
        if move_idx != -1:
            for i in range(move_idx, len(self.args)):
                action = parse_ucci_move(self.args[i])
                if not self.is_red_turn:
                    action = flip_move(action)
                self.history.append(action)
                self.state = self.env.step(self.state, action)
                self.is_red_turn = not self.is_red_turn
                self.turns += 1
                self.history.append(self.state)

    def cmd_fen(self):
        self.args.insert(0, 'fen')
        self.cmd_position()

    # ---- uci.py:177-227
    def cmd_go(self):
        if not self.is_ready:
            return
        self.start_time = time()
        self.t = None
        depth = None
        infinite = True
        self.remain_time = None
        pipes = None
        if self.use_pipes:
            if self.pipes_factory is not None:
                pipes = self.pipes_factory()
            else:
                self.model.close_pipes()
                pipes = self.model.get_pipes(need_reload=False)
        self.player = CChessPlayer(self.config, search_tree=None, pipes=pipes, enable_resign=False, debugging=True, uci=True,
                                   use_history=self.use_history, side=self.turns % 2, lib=self.lib, device=self.device,
                                   weights=None if pipes is not None else self.model.torch_weights(),
                                   infinite_capacity=self.infinite_capacity)
        self.player.info_stream = self.stdout
        for i in range(len(self.args)):
            if self.args[i] == 'depth':
                depth = int(self.args[i + 1]) * 100
                infinite = False
            if self.args[i] == 'movetime' or self.args[i] == 'time':
                self.remain_time = int(self.args[i + 1]) / 1000
            if self.args[i] == 'infinite':
                infinite = True
            if self.args[i] == 'wtime' and self.is_red_turn:
                self.remain_time = int(self.args[i + 1]) / 1000
                depth = 3000
                infinite = False
            if self.args[i] == 'btime' and not self.is_red_turn:
                self.remain_time = int(self.args[i + 1]) / 1000
                depth = 3000
                infinite = False
        self.search_worker = Thread(target=self.search_action, args=(self.player, depth, infinite), daemon=True)
        self.search_worker.start()
        if self.remain_time:
            self.t = Timer(max(self.remain_time - 0.01, 0.0), self.cmd_stop)
            self.t.start()

    # ---- uci.py:229-243
    def cmd_stop(self):
        if not self.is_ready:
            return
        player, self.player = self.player, None
        if player is None:
            return
        no_act = None
        if self.state in self.history[:-1]:
            no_act = []
            for i in range(len(self.history) - 1):
                if self.history[i] == self.state:
                    no_act.append(self.history[i + 1])
        got = player.close_and_return_action(self.state, self.turns, no_act)
        if got is not None:
            self.info_best_move(player, *got)
        self._release(player)

    def cmd_quit(self):
        if self.t:
            self.t.cancel()
        return "quit"

    # ---- uci.py:245-263
    def load_model(self):
        if self.model is None:
            from .model import CChessModel
            self.model = CChessModel(self.config)
            rc = self.config.resource
            if not self.model.load(rc.model_best_config_path, rc.model_best_weight_path):
                self.model.build()
        return bool(self.model.use_history)

    # ---- uci.py:265-291
    def search_action(self, player, depth, infinite):
        no_act = None
        res = self.env.done(self.state, need_check=True)
        check = res[3] if len(res) > 3 else False
        if not check and self.state in self.history[:-1]:
            no_act = []
            for i in range(len(self.history) - 1):
                if self.history[i] == self.state:
                    if self.env.will_check_or_catch(self.state, self.history[i + 1]):
                        no_act.append(self.history[i + 1])
        action, _ = player.action(self.state, self.turns, no_act=no_act, depth=depth, infinite=infinite, hist=self.history)
        if self.player is not player:          # `stop` answered meanwhile (close_and_return_action)
            return
        self.player = None
        if self.t:
            self.t.cancel()
        if action is not None:
            _, value = player.debug[self.state]
            self.info_best_move(player, action, value, player.done_tasks // 100)
        self._release(player)

    def _release(self, player):
        player.close(wait=False)
        if self.use_pipes and self.pipes_factory is None:
            self.model.close_pipes()

    # ---- uci.py:293-327
    def info_best_move(self, player, action, value, depth):
        end_time = time()
        if not self.is_red_turn:
            value = -value
        score = int(value * 1000)
        duration = max(end_time - self.start_time, 1e-9)
        nps = int(depth * 100 / duration) * 1000
        self._print(f"info depth {depth} score {score} time {int(duration * 1000)} nps {nps}")
        # the most visited reply, if the position after `action` is in the tree (first maximum, uci.py:305-311)
        ponder = None
        child = player.engine_child_stats(self.env.step(self.state, action))
        cnt = 0
        for mov, n in child:
            if n > cnt:
                ponder, cnt = mov, n
        if not self.is_red_turn:
            action = flip_move(action)
        output = f"bestmove {to_uci_move(action)}"
        if ponder:
            if self.is_red_turn:
                ponder = flip_move(ponder)
            output += f" ponder {to_uci_move(ponder)}"
        self._print(output)


def main(config):
    UCI(config).main()
