"""oracle/arena.py — CPU restatement of the evaluator's game loop.  TEST INFRASTRUCTURE.

Restates worker/evaluator.py:147-250 (EvaluateWorker.start_game): two players with SEPARATE search trees alternate, the
repetition rule has no be_catched exemption and raises the temperature on any repetition (:173-193), resignation is off
(:157-160).  `draws_for(slot)` supplies the move sampler of the engine slot that plays the ply, so the device arena
(csrc/cz_selfplay.cuh, E.arena) can be compared move for move.  Score bookkeeping of EvaluateWorker.start (:93-145) is
`score_for_next_generation`.

Pinned: replays whole games of the REAL, unmodified EvaluateWorker.start_game move for move when the random decisions come
from the reference's own generators (tests/golden/games_k1.json.gz, tests/test_games_golden.py).
"""
from . import senv
from .player import OraclePlayer


def play_arena_game(pc, evaluate0, evaluate1, idx, draws_for, m_games, max_game_length=100, env=senv, max_plies_guard=1000,
                    playouts=None):
    """idx = running game index; player idx % 2 is red (evaluator.py:163-170).  Returns dict(moves, value_red, turns, flags).
    playouts: (lo, hi) -> the game first draws `randint(lo, hi) * 100` simulations per move for BOTH players
    (evaluator.py:153-154 rebinds config.play.simulation_num_per_move before the players are created), through
    `draws_for(slot).playouts(lo, hi)`; None = pc.simulation_num_per_move as given."""
    sims_game = None
    if playouts is not None:
        import copy
        sims_game = draws_for(idx % m_games).playouts(*playouts)
        pc = copy.copy(pc)
        pc.simulation_num_per_move = sims_game
    # the reference draws its Dirichlet sample even when eps == 0 (player.py:304): keep that when replaying its RNG
    quiet = (lambda n: 0.0) if (pc.noise_eps == 0 and not hasattr(draws_for(0), "choose_with_player")) else None
    players = [OraclePlayer(pc, evaluate0, env=env, noise=quiet), OraclePlayer(pc, evaluate1, env=env, noise=quiet)]
    i = idx % m_games
    state = env.INIT_STATE
    history = [state]
    value, turns, game_over, final_move = 0, 0, False, None
    no_eat_count, check = 0, False
    flags = 0
    while not game_over and turns < max_plies_guard:
        no_act, increase_temp = None, False
        if not check and state in history[:-1]:
            no_act, increase_temp, free_move = [], True, 0
            for k in range(len(history) - 1):
                if history[k] == state:
                    if env.will_check_or_catch(state, history[k + 1]):
                        no_act.append(history[k + 1])
                    else:
                        free_move += 1
                        if free_move >= 3:
                            game_over, value = True, 0
                            flags |= 2
                            break
        if game_over:
            break
        p = (idx + turns) % 2                                  # red = player idx % 2 moves on even plies
        player = players[p]
        player.search(state, no_act, increase_temp)
        node = player.tree[state]
        draws = draws_for(i + p * m_games)
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
            if not game_over and not env.has_attack_chessman(state):
                game_over, value = True, 0
                flags |= 2
    if final_move:
        history.append(final_move)
        state = env.step(state, final_move)
        turns += 1
        value = -value
        history.append(state)
    if turns % 2 == 1:
        value = -value
    return {"moves": [history[2 * k + 1] for k in range(turns)], "value_red": value, "turns": turns, "flags": flags,
            "playouts": sims_game}


def score_for_next_generation(value_red, idx):
    """evaluator.py:127-137: score of the next-generation model (player 1); best model (player 0) is red when idx is even."""
    score = 0 if value_red == -1 else (1 if value_red == 1 else 0.5)
    return 1 - score if idx % 2 == 0 else score
