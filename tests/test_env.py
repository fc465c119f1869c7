"""Rules-engine parity: kernels == golden vectors produced by the real reference."""
import pytest

from tests import env_checks


def test_oracle_env_matches_golden(golden_env):
    from oracle import senv as o
    assert golden_env["n_labels"] == len(o.ActionLabelsRed) == 2086
    assert o.ActionLabelsRed[:12] == golden_env["labels_first"] and o.ActionLabelsRed[-5:] == golden_env["labels_last"]
    import hashlib
    assert hashlib.sha256("".join(o.ActionLabelsRed).encode()).hexdigest() == golden_env["labels_sha"]
    for r in golden_env["rows"]:
        s = r["state"]
        assert o.get_legal_moves(s) == r["moves"]
        assert list(o.done(s, need_check=True)) == r["done"]
        assert o.state_to_planes(s).reshape(-1).nonzero()[0].tolist() == r["plane_idx"]
        assert o.fliped_state(s) == r["flip"] and o.has_attack_chessman(s) == r["attack"]
        if "move" in r:
            assert o.new_step(s, r["move"]) == (r["next"], r["no_eat"])
            assert o.will_check_or_catch(s, r["move"]) == r["wcc"] and o.be_catched(s, r["move"]) == r["bc"]


def test_survey_golden_vectors():
    from oracle import senv as o
    """The vectors SURVEY.md §4 recovered from the reference's print-style scripts (test.py:112-203), as assertions."""
    init_moves = ("0001 0002 1022 1002 2042 2002 3041 4041 5041 6082 6042 7082 7062 8081 8082 1202 1222 1232 1242 1252 1262 "
                  "1211 1213 1214 1215 1216 1219 7222 7232 7242 7252 7262 7282 7271 7273 7274 7275 7276 7279 0304 2324 4344 "
                  "6364 8384").split()
    assert o.get_legal_moves(o.INIT_STATE) == init_moves
    p = o.state_to_planes(o.INIT_STATE)
    assert p.shape == (14, 10, 9) and p.sum() == 32 and p.sum(axis=(1, 2)).tolist() == [5, 2, 2, 2, 2, 2, 1, 5, 2, 2, 2, 2, 2, 1]
    assert p[6, 9, 4] == 1 and p[13, 0, 4] == 1
    labels = o.ActionLabelsRed
    assert len(labels) == len(set(labels)) == 2086 and labels[:2] == ['0010', '0020'] and labels[-5:] == ['8769', '0725', '4725', '4765', '8765']
    s = '4s4/9/4e4/p8/2e2R2p/P5E2/8P/9/9/4S1E2'
    lm = o.get_legal_moves(s)
    assert len(lm) == 24 and lm[:12] == '4050 4041 4030 6082 6042 8384 0405 6442 6482 5535 5545 5565'.split() and lm[-3:] == ['5559', '5525', '5585']
    assert o.state_to_fen(o.step(o.INIT_STATE, '0001'), 1) == 'rnbakabnr/9/1c5c1/p1p1p1p1p/9/9/P1P1P1P1P/1C5C1/R8/1NBAKABNR b - - 0 1'
    assert o.parse_ucci_move('b7b0') == '1710' and o.flip_move('1710') == '7279'


def test_emul_env_golden(emul_env, golden_env):
    env_checks.check_against_rows(emul_env, golden_env["rows"])
    env_checks.check_keys(emul_env, golden_env["rows"])


def test_emul_env_single_api(emul_env):
    env_checks.check_single_api(emul_env)


def test_emul_env_extreme_positions(emul_env):
    env_checks.check_extreme_positions(emul_env)


def test_emul_env_random_boards(emul_env):
    env_checks.check_random_boards(emul_env)


@pytest.mark.gpu
def test_cuda_env_extreme_positions(cuda_env):
    env_checks.check_extreme_positions(cuda_env)
    env_checks.check_random_boards(cuda_env, n=2000, seed=13)


@pytest.mark.gpu
def test_cuda_env_golden(cuda_env, golden_env):
    env_checks.check_against_rows(cuda_env, golden_env["rows"])
    env_checks.check_keys(cuda_env, golden_env["rows"])


@pytest.mark.gpu
def test_cuda_env_single_api(cuda_env):
    env_checks.check_single_api(cuda_env)


def test_emul_env_playout_sweep(emul_env):
    """The bulk sweep of tests/env_sweep.py on the emulator build (small: the emulator runs lanes one after another)."""
    from tests.env_sweep import check_sweep
    st = check_sweep(emul_env, 4000, seed=7, procs=4)
    assert st["terminal"] > 20 and st["wcc"] > 100 and st["bc"] > 100


@pytest.mark.gpu
def test_cuda_env_playout_sweep_1e5(cuda_env):
    """>= 1e5 random-playout positions, every rules kernel bit for bit against the oracle."""
    from tests.env_sweep import check_sweep
    st = check_sweep(cuda_env, 120000, seed=2024)
    print("rules sweep:", st)
    assert st["positions"] == 120000 and st["terminal"] > 1000 and st["checks"] > 500 and st["wcc"] > 5000 and st["bc"] > 5000
