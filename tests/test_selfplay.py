"""On-device game loop (cz_play_move) vs the restated reference loop (worker/self_play.py:95-212), move for move."""
import numpy as np
import pytest

from cczero_b200.engine import Engine
from oracle import player as op
from oracle import selfplay as osp
from oracle import senv as osenv
from tests.search_checks import eval_planes


def run_device_games(lib, device, n_games, sims, k, seed, want, max_game_length, tau_decay, use_history=False):
    eng = Engine(lib, device, n_games=n_games, sims_per_move=sims, leaves_per_round=k, noise_mode=1, noise_eps=0.0,
                 c_puct=1.5, tau_decay_rate=tau_decay, max_game_length=max_game_length, resign_threshold=-0.6,
                 enable_resign_rate=0.5, min_resign_turn=4, seed=seed, max_nodes_per_game=sims * 2 * max_game_length + 64,
                 use_history=use_history)
    eng.reset()
    recs = []
    for _ in range(4 * max_game_length * (want // n_games + 2)):
        eng.search_external(eval_planes, None)
        if eng.play_move():
            recs += eng.drain_records()
        if len(recs) >= want:
            break
    assert int(eng.counters()[4]) == 0          # no tree reset happened (pools were large enough)
    eng.close()
    return recs


def check_selfplay(lib, device, n_games=3, sims=20, k=4, seed=11, want=6, max_game_length=25, tau_decay=0.9, use_history=False):
    recs = run_device_games(lib, device, n_games, sims, k, seed, want, max_game_length, tau_decay, use_history)
    assert len(recs) >= want
    label_of = {m: i for i, m in enumerate(osenv.ActionLabelsRed)}
    pc = op.PlayConfig(simulation_num_per_move=sims, search_threads=k, c_puct=1.5, noise_eps=0.0, dirichlet_alpha=0.2,
                       tau_decay_rate=tau_decay, virtual_loss=3, resign_threshold=-0.6, min_resign_turn=4)
    kinds = set()
    for r in recs:
        slot, started = r["game_index"] % n_games, r["game_index"] // n_games
        ref = osp.play_game(pc, op.fake_evaluate_states_hist if use_history else op.fake_evaluate_states,
                            osp.DeviceDraws(seed, 0, slot, started, label_of),
                            max_game_length=max_game_length, enable_resign_rate=0.5, use_history=use_history)
        assert r["moves"] == ref["moves"], (r["game_index"], r["moves"], ref["moves"])
        assert r["value_red"] == ref["value_red"] and r["n_plies"] == ref["turns"]
        assert (r["flags"] & 3) == ref["flags"] and bool(r["flags"] & 4) == (not ref["store"])
        kinds.add((r["flags"] & 3, r["value_red"]))
    return kinds


def test_emul_selfplay_matches_restated_game_loop(emul_lib):
    kinds = check_selfplay(emul_lib, "cpu")
    assert len(kinds) >= 1


def test_emul_selfplay_other_configurations(emul_lib):
    """More shapes of the device game loop against the restated one: K = 1 and K > simulations, no temperature decay
    (always arg-max), long temperature schedule, short and longer games."""
    check_selfplay(emul_lib, "cpu", n_games=2, sims=12, k=1, seed=3, want=3, max_game_length=14, tau_decay=0.0)
    check_selfplay(emul_lib, "cpu", n_games=2, sims=10, k=16, seed=4, want=2, max_game_length=18, tau_decay=0.99)
    check_selfplay(emul_lib, "cpu", n_games=1, sims=30, k=8, seed=8, want=2, max_game_length=30, tau_decay=0.5)


def test_emul_selfplay_with_history_planes(emul_lib):
    """use_history=True: the device game loop feeds 28-plane leaves (path history only, self_play.py:124)."""
    check_selfplay(emul_lib, "cpu", n_games=2, want=3, seed=5, use_history=True)


def test_play_data_format():
    from cczero_b200.records import record_to_play_data
    d = record_to_play_data({"moves": ["7747", "7062", "1219"], "value_red": -1})
    assert d == [osenv.INIT_STATE, ["7747", -1], ["7062", 1], ["1219", -1]]


@pytest.mark.gpu
def test_cuda_selfplay_matches_restated_game_loop(cuda_lib):
    check_selfplay(cuda_lib, "cuda", n_games=4, want=8)


@pytest.mark.gpu
def test_cuda_selfplay_other_configurations(cuda_lib):
    check_selfplay(cuda_lib, "cuda", n_games=8, sims=12, k=1, seed=3, want=8, max_game_length=14, tau_decay=0.0)
    check_selfplay(cuda_lib, "cuda", n_games=8, sims=10, k=16, seed=4, want=8, max_game_length=18, tau_decay=0.99)
    check_selfplay(cuda_lib, "cuda", n_games=4, sims=60, k=8, seed=8, want=4, max_game_length=60, tau_decay=0.5)


@pytest.mark.gpu
def test_cuda_selfplay_with_history_planes(cuda_lib):
    check_selfplay(cuda_lib, "cuda", n_games=4, want=6, seed=5, use_history=True)


def test_emul_expanding_data_matches_reference_layout(emul_env):
    """records.expanding_data == optimize.py:234-281 on a replayed game (planes, one-hot policy, alternating value)."""
    from cczero_b200.records import expanding_data, record_to_play_data
    rng = np.random.RandomState(5)
    s, moves = osenv.INIT_STATE, []
    for _ in range(30):
        lm = osenv.get_legal_moves(s)
        m = lm[rng.randint(len(lm))]
        moves.append(m)
        s = osenv.step(s, m)
    data = record_to_play_data({"moves": moves, "value_red": 1})
    planes, policy, value = expanding_data(data, emul_env)
    assert planes.shape == (30, 14, 10, 9) and policy.shape == (30, 2086) and value.shape == (30,)
    s = osenv.INIT_STATE
    for i, m in enumerate(moves):
        assert (planes[i] == osenv.state_to_planes(s)).all()
        assert policy[i].sum() == 1 and policy[i, osenv.ActionLabelsRed.index(m)] == 1
        assert value[i] == (1 if i % 2 == 0 else -1)
        s = osenv.step(s, m)
    # use_history (optimize.py:240-267): planes 14-27 of sample i = position i-2
    planes28, policy28, _ = expanding_data(data, emul_env, use_history=True)
    assert planes28.shape == (30, 28, 10, 9) and (policy28 == policy).all()
    s, hist = osenv.INIT_STATE, [osenv.INIT_STATE]
    for i, m in enumerate(moves):
        assert (planes28[i] == osenv.state_history_to_planes(s, hist[0:2 * i + 1])).all()
        s = osenv.step(s, m)
        hist += [m, s]


def test_flip_policy_matches_reference_definition(emul_env):
    """lookup_tables.py:134-141: Unflipped_index = [ActionLabelsRed.index(x) for x in ActionLabelsBlack]."""
    from cczero_b200.records import build_policy, flip_policy
    red = osenv.ActionLabelsRed
    black = [osenv.flip_move(m) for m in red]
    pol = np.random.RandomState(1).rand(len(red))
    want = np.asarray([pol[red.index(x)] for x in black])
    assert (flip_policy(pol, emul_env) == want).all()
    p = build_policy("7747", True, emul_env)
    assert sum(p) == 1 and p[red.index(osenv.flip_move("7747"))] == 1


# ---- bounded runs, the record ring, file batching (host logic + device bookkeeping on the emulator)
def _engine(lib, device, n_games, sims=6, k=4, max_game_length=3, **kw):
    return Engine(lib, device, n_games=n_games, sims_per_move=sims, leaves_per_round=k, noise_mode=1, noise_eps=0.0,
                  tau_decay_rate=0.9, max_game_length=max_game_length, enable_resign_rate=0.0, seed=5,
                  max_nodes_per_game=256, **kw)


def _play_all(eng, guard=400):
    recs = []
    for _ in range(guard):
        eng.search_external(eval_planes, None)
        eng.play_move()
        recs += eng.drain_records()
        if not eng.any_active():
            break
    return recs


def test_emul_game_quota_plays_each_index_once(emul_lib):
    """cz_config.game_quota: exactly the games 0 .. quota-1 are played, each to its end, then every slot retires."""
    eng = _engine(emul_lib, "cpu", n_games=4, game_quota=10)
    recs = _play_all(eng)
    assert sorted(r["game_index"] for r in recs) == list(range(10))
    assert all(r["n_plies"] <= 7 for r in recs) and sum(r["n_plies"] == 6 for r in recs) >= 8   # max_game_length = 3: 6-ply draws
    eng.close()
    eng = _engine(emul_lib, "cpu", n_games=4, game_quota=3)          # more slots than games: slot 3 never plays
    assert sorted(r["game_index"] for r in _play_all(eng)) == [0, 1, 2]
    eng.close()


def test_config_rejects_short_record_rows(emul_lib):
    """max_plies must cover 2*max_game_length (the history / record rows are written up to that index)."""
    import ctypes as C
    from cczero_b200.lib import CzConfig, CzError
    eng = _engine(emul_lib, "cpu", n_games=1)
    cfg = CzConfig.from_buffer_copy(eng.cfg)
    eng.close()
    cfg.max_plies = 2 * cfg.max_game_length - 1
    n = C.c_uint64(0)
    with pytest.raises(CzError, match="max_plies"):
        emul_lib.call("cz_workspace_bytes", C.byref(cfg), C.byref(n))


def test_emul_record_ring_is_never_silently_overrun(emul_lib):
    """cz_selfplay returns early while the ring can still take a ply's worth of finished games; driving cz_play_move
    without draining past the ring capacity fails loudly and counts the dropped records."""
    from cczero_b200.lib import CzError
    eng = _engine(emul_lib, "cpu", n_games=40, sims=2, k=2, max_game_length=1)       # every game ends after 2 plies; ring = 80
    with pytest.raises(CzError, match="dropped"):
        for _ in range(8):                                                            # 40 records per 2 plies, never drained
            eng.search_external(eval_planes, None)
            eng.play_move()
    assert int(eng.counters()[3]) > 0                                                 # counted, not silent
    assert len(eng.drain_records()) == 80                                             # what fitted is intact
    eng.close()


def test_play_data_files_with_several_games_per_file(tmp_path):
    """with nb_game_in_file = 5 (configs/normal.py) a file must be written every 5 stored games."""
    from types import SimpleNamespace
    from cczero_b200.self_play import SelfPlayWorker
    w = SelfPlayWorker.__new__(SelfPlayWorker)
    w.config = SimpleNamespace(play_data=SimpleNamespace(nb_game_in_file=5),
                               resource=SimpleNamespace(play_data_dir=str(tmp_path), play_data_filename_tmpl="play_%s.json"))
    w.buffer, w.games_written, w.games_stored, w.pending, w.pid = [], 0, 0, [], 0
    recs = [{"moves": ["7747", "7062"], "value_red": 1, "flags": 0, "n_plies": 2, "game_index": i} for i in range(12)]
    recs[3]["flags"] = 4                                                              # not stored (short-game lottery)
    w.pending = list(recs)
    w.engine = None
    out = w.play_games(12)
    assert len(out) == 12 and w.games_stored == 11 and w.games_written == 2
    import glob, json
    files = sorted(glob.glob(str(tmp_path / "play_*.json")))
    assert len(files) == 2
    assert all(len(json.load(open(f))) == 5 * 3 for f in files)                      # 5 games x (init state + 2 moves)
    assert len(w.buffer) == 3                                                         # the 11th game waits for the next file
