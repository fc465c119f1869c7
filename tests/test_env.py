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
