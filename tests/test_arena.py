"""Evaluator arena (cz_config.arena): two players with separate trees per game, evaluator draw rules, colour alternation —
move for move against the restated worker/evaluator.py game loop."""
import pytest

from cczero_b200.engine import Engine
from oracle import arena as oarena
from oracle import player as op
from oracle import selfplay as osp
from oracle import senv as osenv
from tests.search_checks import eval_planes


def check_arena(lib, device, m_games=3, sims=20, k=4, seed=31, want=6, max_game_length=20, playouts=None):
    """playouts=(lo, hi): every game draws its own simulations per move = randint(lo, hi) * 100 from its Philox stream
    (evaluator.py:153-154); the restated loop takes the same draw through DeviceDraws.playouts."""
    top = playouts[1] * 100 if playouts else sims
    eng = Engine(lib, device, n_games=2 * m_games, sims_per_move=sims, leaves_per_round=k, noise_mode=1, noise_eps=0.0,
                 c_puct=1.0, tau_decay_rate=0.0, max_game_length=max_game_length, enable_resign_rate=0.0, seed=seed,
                 max_nodes_per_game=top * 2 * max_game_length + 64, arena=True, playouts=playouts)
    eng.reset()
    recs = []
    for _ in range(2 * max_game_length * (want // m_games + 2) + 8):
        eng.search_external(eval_planes, None)
        if eng.play_move():
            recs += eng.drain_records()
        if len(recs) >= want:
            break
    assert len(recs) >= want and int(eng.counters()[6]) == 0 and int(eng.counters()[4]) == 0
    eng.close()
    label_of = {m: i for i, m in enumerate(osenv.ActionLabelsRed)}
    pc = op.PlayConfig(simulation_num_per_move=sims, search_threads=k, c_puct=1.0, noise_eps=0.0, tau_decay_rate=0.0, virtual_loss=3)
    drawn = set()
    for r in recs:
        idx = r["game_index"]
        started = idx // m_games
        ref = oarena.play_arena_game(pc, op.fake_evaluate_states, op.fake_evaluate_states, idx,
                                     lambda slot: osp.DeviceDraws(seed, 0, slot, started, label_of), m_games,
                                     max_game_length=max_game_length, playouts=playouts)
        if playouts:
            drawn.add(ref["playouts"])
        assert r["moves"] == ref["moves"], (idx, r["moves"], ref["moves"])
        assert r["value_red"] == ref["value_red"] and r["n_plies"] == ref["turns"] and (r["flags"] & 3) == ref["flags"]
    if playouts:
        assert len(drawn) >= 2 and drawn <= {100 * v for v in range(playouts[0], playouts[1] + 1)}, drawn
    assert oarena.score_for_next_generation(1, 0) == 0 and oarena.score_for_next_generation(1, 1) == 1
    assert oarena.score_for_next_generation(0, 4) == 0.5


def test_emul_arena_matches_restated_evaluator_loop(emul_lib):
    check_arena(emul_lib, "cpu")


def test_emul_arena_per_game_playouts(emul_lib):
    """Per-game `randint(lo, hi) * 100` simulations per move, both player slots of a game on the same value."""
    check_arena(emul_lib, "cpu", m_games=3, k=8, seed=5, want=6, max_game_length=4, playouts=(1, 3))


@pytest.mark.gpu
def test_cuda_arena_matches_restated_evaluator_loop(cuda_lib):
    check_arena(cuda_lib, "cuda", m_games=4, want=8)
    check_arena(cuda_lib, "cuda", m_games=4, k=8, seed=5, want=8, max_game_length=6, playouts=(1, 4))
