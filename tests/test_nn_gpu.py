"""Network forward parity: tensor-core pipeline vs the fp32 PyTorch restatement of agent/model.py
(tolerance 1e-3 on policy probabilities and value, north_star)."""
import numpy as np
import pytest
import torch

from oracle import model as om
from oracle import senv as osenv
from tests.search_checks import midgame_states

pytestmark = pytest.mark.gpu


def _engine(cuda_lib, filters, blocks, batch, fp32_skip=None, use_history=False):
    from cczero_b200.engine import Engine
    return Engine(cuda_lib, "cuda", n_games=batch, sims_per_move=8, leaves_per_round=1, nn_filters=filters,
                  nn_blocks=blocks, nn_value_fc=256, nn_fp32_skip=fp32_skip, use_history=use_history)


# Tolerance 1e-3 on policy probabilities and value (north_star), asserted on
#   * Keras-default-initialised nets (what `run.py self --new` builds, agent/model.py:32-66) of every BASELINE size, and
#   * nets with mildly perturbed BatchNorm statistics / biases (spread 0.3) so that a folding bug cannot hide.
# Measured (tools/nn_error_report.py, profiles/): policy <= 1.4e-4 everywhere; value <= 6e-4 with the default precision
# policy (skip stream fp16 up to 10 blocks, fp32 beyond).  Strongly perturbed random BN statistics (spread 1.0) make a
# 10-20 block random net amplify ANY operand rounding several-fold (value deviations up to 3.4e-3 were measured even
# with the fp32 skip stream); those nets are bounded separately at 1e-2 as a gross-error check.
@pytest.mark.parametrize("filters,blocks,trained,spread,fp32_skip", [
    (128, 7, False, 0, None), (256, 7, False, 0, None), (192, 10, False, 0, None), (256, 20, False, 0, None),
    (128, 7, True, 0.3, None), (256, 3, True, 1.0, None), (192, 10, True, 0.3, None), (256, 20, True, 0.1, None),
    (192, 2, True, 1.0, True)])
def test_forward_matches_fp32_restatement(cuda_lib, cuda_env, filters, blocks, trained, spread, fp32_skip):
    w = om.init_weights(filters, blocks, 256, seed=filters + blocks, trained_like=trained, spread=spread)
    states = [osenv.INIT_STATE] + midgame_states(40, 3, lo=1, hi=120)
    planes = np.stack([osenv.state_to_planes(s) for s in states])
    ref_p, ref_v = om.forward(w, planes, blocks)
    eng = _engine(cuda_lib, filters, blocks, 64, fp32_skip)
    eng.set_weights({k: torch.as_tensor(v) for k, v in w.items()})
    pol, val = eng.nn_forward_planes(torch.as_tensor(planes).cuda())
    pol2, val2 = eng.nn_forward_boards(cuda_env.boards_from_states(states))
    torch.cuda.synchronize()
    assert torch.equal(pol, pol2) and torch.equal(val, val2)          # fused plane encoding == explicit planes
    pol, val = pol.cpu().numpy(), val.cpu().numpy()
    assert np.isfinite(pol).all() and np.isfinite(val).all()
    assert np.abs(pol.sum(1) - 1).max() < 1e-4
    assert np.abs(pol - ref_p).max() < 1e-3, np.abs(pol - ref_p).max()
    assert np.abs(val - ref_v).max() < 1e-3, np.abs(val - ref_v).max()
    assert (pol.argmax(1) == ref_p.argmax(1)).mean() > 0.9            # the ordering of the top moves is what the search consumes
    eng.close()


@pytest.mark.parametrize("filters,blocks", [(192, 10), (256, 20)])
def test_forward_on_ill_conditioned_random_nets(cuda_lib, cuda_env, filters, blocks):
    """Gross-error bound (1e-2) on deep random nets with strongly perturbed BN statistics (see the comment above)."""
    w = om.init_weights(filters, blocks, 256, seed=filters + blocks, trained_like=True, spread=1.0)
    states = [osenv.INIT_STATE] + midgame_states(40, 3, lo=1, hi=120)
    ref_p, ref_v = om.forward(w, np.stack([osenv.state_to_planes(s) for s in states]), blocks)
    eng = _engine(cuda_lib, filters, blocks, 64)
    eng.set_weights({k: torch.as_tensor(v) for k, v in w.items()})
    pol, val = eng.nn_forward_boards(cuda_env.boards_from_states(states))
    assert np.abs(pol.cpu().numpy() - ref_p).max() < 1e-2 and np.abs(val.cpu().numpy() - ref_v).max() < 1e-2
    eng.close()


def test_forward_chunks_and_batch_of_one(cuda_lib, cuda_env):
    w = om.init_weights(128, 2, 256, seed=1, trained_like=True)
    states = midgame_states(9, 5)
    eng = _engine(cuda_lib, 128, 2, 4)       # max batch 4 -> 9 positions run as 3 chunks
    eng.set_weights({k: torch.as_tensor(v) for k, v in w.items()})
    b = cuda_env.boards_from_states(states)
    pol, val = eng.nn_forward_boards(b)
    p1, v1 = eng.nn_forward_boards(b[4:5])
    torch.cuda.synchronize()
    assert torch.allclose(pol[4], p1[0], atol=1e-6) and torch.allclose(val[4], v1[0], atol=1e-6)
    ref_p, ref_v = om.forward(w, np.stack([osenv.state_to_planes(s) for s in states]), 2)
    assert np.abs(pol.cpu().numpy() - ref_p).max() < 1e-3 and np.abs(val.cpu().numpy() - ref_v).max() < 1e-3
    eng.close()


@pytest.mark.parametrize("filters,blocks,fp32_skip", [(128, 7, None), (192, 10, None), (256, 20, None), (256, 3, False)])
def test_small_batch_tiles_equal_full_width_tiles(cuda_lib, cuda_env, filters, blocks, fp32_skip):
    """A few positions (one game's leaves: UCI / play_games) run the residual convs as 64-column tiles spread over many CTA pairs
    (cz_nn.cu use_n_split); the same positions inside a batch of 1024 run the full-width tiles.  The K order of every output is the
    same, so policy and value agree bit for bit, and both keep the 1e-3 bound."""
    w = om.init_weights(filters, blocks, 256, seed=3, trained_like=True, spread=0.1)
    states = [osenv.INIT_STATE] + midgame_states(9, 11, lo=1, hi=100)
    boards = cuda_env.boards_from_states(states)
    small = _engine(cuda_lib, filters, blocks, 16, fp32_skip)
    big = _engine(cuda_lib, filters, blocks, 1024, fp32_skip)
    for e in (small, big):
        e.set_weights({k: torch.as_tensor(v) for k, v in w.items()})
    p_s, v_s = small.nn_forward_boards(boards)
    reps = (1024 + len(states) - 1) // len(states)
    p_b, v_b = big.nn_forward_boards(boards.repeat(reps, 1)[:1024])
    torch.cuda.synchronize()
    n = len(states)
    assert torch.equal(p_s, p_b[:n]) and torch.equal(v_s, v_b[:n])
    assert torch.equal(p_b[:n], p_b[n:2 * n])
    ref_p, ref_v = om.forward(w, np.stack([osenv.state_to_planes(s) for s in states]), blocks)
    assert np.abs(p_s.cpu().numpy() - ref_p).max() < 1e-3 and np.abs(v_s.cpu().numpy() - ref_v).max() < 1e-3
    small.close(); big.close()


def test_cluster4_weight_multicast(cuda_lib, cuda_env, monkeypatch):
    """Experiment kept behind CZ_CLUSTER4=1 (cz_igemm3.cuh, PAIRS = 2): the two CTA pairs of a 4-CTA cluster share every weight
    stage of the 256-wide conv by TMA multicast.  Same K order per output: policy and value bit for bit those of the default launch,
    odd numbers of tiles included (the second pair then walks an all-out-of-bounds tile)."""
    w = om.init_weights(256, 3, 256, seed=9, trained_like=True, spread=0.1)
    states = [osenv.INIT_STATE] + midgame_states(12, 17, lo=1, hi=100)
    boards = cuda_env.boards_from_states(states)
    out = {}
    for mode in ("0", "1"):                                   # the switch is read when a network runtime is created
        monkeypatch.setenv("CZ_CLUSTER4", mode)
        for batch in (302, 1024, 1027):      # 302 boards = 107 pair-tiles: odd
            eng = _engine(cuda_lib, 256, 3, batch, False)     # fp16 skip stream: both convs of a block run the TMA-epilogue kernel
            eng.set_weights({k: torch.as_tensor(v) for k, v in w.items()})
            reps = (batch + len(states) - 1) // len(states)
            p, v = eng.nn_forward_boards(boards.repeat(reps, 1)[:batch])
            torch.cuda.synchronize()
            out[(mode, batch)] = (p.clone(), v.clone())
            eng.close()
    monkeypatch.setenv("CZ_CLUSTER4", "0")
    _engine(cuda_lib, 64, 1, 4).close()                       # leave the process-wide switch off for the tests that follow
    for batch in (302, 1024, 1027):      # 302 boards = 107 pair-tiles: odd
        assert torch.equal(out[("0", batch)][0], out[("1", batch)][0]) and torch.equal(out[("0", batch)][1], out[("1", batch)][1])
    ref_p, ref_v = om.forward(w, np.stack([osenv.state_to_planes(s) for s in states]), 3)
    n = len(states)
    assert np.abs(out[("1", 1024)][0][:n].cpu().numpy() - ref_p).max() < 1e-3 and np.abs(out[("1", 1024)][1][:n].cpu().numpy() - ref_v).max() < 1e-3


@pytest.mark.parametrize("filters,blocks,trained", [(128, 7, False), (192, 4, True)])
def test_forward_28_planes_with_history(cuda_lib, cuda_env, filters, blocks, trained):
    """use_history networks (data/model/model_128_l1_config.json: Input (28,10,9)): planes 14-27 = the position two plies
    earlier or zero (static_env.py:158-194); the first convolution gathers from (board, history board) pairs."""
    from tests.search_checks import game_history
    w = om.init_weights(filters, blocks, 256, seed=5, trained_like=trained, spread=0.3, in_planes=28)
    hists = [game_history(n, 100 + n) for n in (1, 2, 3, 9, 24, 40, 61)] + [None]
    states = [h[-1] for h in hists[:-1]] + [osenv.INIT_STATE]
    planes = np.stack([osenv.state_history_to_planes(s, h) for s, h in zip(states, hists)])
    assert planes[3, 14:].sum() > 0 and planes[0, 14:].sum() == 0 and planes[-1, 14:].sum() == 0
    ref_p, ref_v = om.forward(w, planes, blocks)
    eng = _engine(cuda_lib, filters, blocks, 8, use_history=True)
    eng.set_weights({k: torch.as_tensor(v) for k, v in w.items()})
    pol, val = eng.nn_forward_planes(torch.as_tensor(planes).cuda())
    pairs = torch.zeros((len(states), 2, 96), dtype=torch.uint8, device="cuda")
    pairs[:, 0] = cuda_env.boards_from_states(states)
    for i, h in enumerate(hists):
        if h and len(h) >= 5:
            pairs[i, 1] = cuda_env.boards_from_states([h[-5]])[0]
    pol2, val2 = eng.nn_forward_boards(pairs.reshape(len(states), 192))
    torch.cuda.synchronize()
    assert torch.equal(pol, pol2) and torch.equal(val, val2)
    assert np.abs(pol.cpu().numpy() - ref_p).max() < 1e-3 and np.abs(val.cpu().numpy() - ref_v).max() < 1e-3
    # the history planes matter: dropping them changes the output
    pairs[:, 1] = 0
    pol3, _ = eng.nn_forward_boards(pairs.reshape(len(states), 192))
    assert not torch.equal(pol3[3], pol[3]) and torch.equal(pol3[0], pol[0])
    eng.close()
    # a 14-plane weight set is refused by a use_history engine and vice versa
    eng = _engine(cuda_lib, filters, blocks, 8, use_history=False)
    with pytest.raises(Exception):
        eng.set_weights({k: torch.as_tensor(v) for k, v in w.items()})
    eng.close()


def test_deep_net_fp16_skip_stream_bound(cuda_lib, cuda_env):
    """Forcing the fp16 skip stream on 20 blocks (the faster, non-default mode) stays within 3e-3 on the value."""
    w = om.init_weights(256, 20, 256, seed=276, trained_like=False)
    states = [osenv.INIT_STATE] + midgame_states(40, 3, lo=1, hi=120)
    ref_p, ref_v = om.forward(w, np.stack([osenv.state_to_planes(s) for s in states]), 20)
    eng = _engine(cuda_lib, 256, 20, 64, fp32_skip=False)
    eng.set_weights({k: torch.as_tensor(v) for k, v in w.items()})
    pol, val = eng.nn_forward_boards(cuda_env.boards_from_states(states))
    assert np.abs(pol.cpu().numpy() - ref_p).max() < 1e-3
    assert np.abs(val.cpu().numpy() - ref_v).max() < 3e-3
    eng.close()


@pytest.mark.parametrize("filters,blocks,pol_c,val_c,in_planes", [
    (128, 7, 2, 4, 14),       # data/model/model_128f.json
    (256, 7, 2, 4, 14),       # data/model/model_256f.json
    (128, 7, 32, 4, 28),      # data/model/model_128_l1_config.json (28-plane history input)
])
def test_legacy_head_widths(cuda_lib, cuda_env, filters, blocks, pol_c, val_c, in_planes):
    """The older configs shipped under the reference's data/model/ have other head widths than agent/model.py:47-61 builds
    (policy 2 or 32 channels, value 4): the engine serves them, within 1e-3 of the fp32 restatement (which
    tests/test_oracle_vs_reference.py pins to those very JSON files through oracle/keras_graph.py)."""
    from cczero_b200.engine import Engine
    w = om.init_weights(filters, blocks, 256, seed=pol_c + val_c, trained_like=True, spread=0.3, in_planes=in_planes,
                        policy_filters=pol_c, value_filters=val_c)
    assert w["policy_out/kernel"].shape == (90 * pol_c, 2086) and w["value_dense/kernel"].shape == (90 * val_c, 256)
    states = [osenv.INIT_STATE] + midgame_states(24, 5, lo=1, hi=100)
    hist = in_planes == 28
    planes = np.stack([osenv.state_history_to_planes(s, [s]) if hist else osenv.state_to_planes(s) for s in states])
    ref_p, ref_v = om.forward(w, planes, blocks)
    eng = Engine(cuda_lib, "cuda", n_games=32, sims_per_move=8, leaves_per_round=1, nn_filters=filters, nn_blocks=blocks,
                 nn_value_fc=256, use_history=hist, nn_policy_channels=pol_c, nn_value_channels=val_c)
    eng.set_weights({k: torch.as_tensor(v) for k, v in w.items()})
    pol, val = eng.nn_forward_planes(torch.as_tensor(planes).cuda())
    pol, val = pol.cpu().numpy(), val.cpu().numpy()
    assert np.abs(pol.sum(1) - 1).max() < 1e-4
    assert np.abs(pol - ref_p).max() < 1e-3, np.abs(pol - ref_p).max()
    assert np.abs(val - ref_v).max() < 1e-3, np.abs(val - ref_v).max()
    # wrong widths are refused loudly, not silently mis-read
    from cczero_b200.lib import CzError
    eng2 = Engine(cuda_lib, "cuda", n_games=8, sims_per_move=8, leaves_per_round=1, nn_filters=filters, nn_blocks=blocks, use_history=hist)
    with pytest.raises(CzError, match="mis-sized"):
        eng2.set_weights({k: torch.as_tensor(v) for k, v in w.items()})
    eng.close()
    eng2.close()
