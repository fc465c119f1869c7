"""Host logic of the CChessModelAPI drop-in (reference: agent/api.py:37-74) without a GPU: the batching / scatter loop
served by a stand-in engine.  The network itself is covered by the -m gpu tier."""
import threading
import time
from types import SimpleNamespace

import numpy as np
import torch

from oracle import player as op
from oracle import senv as osenv


class StubEngine:
    """nn_forward_planes of the real Engine, computed by the deterministic pseudo-network."""
    lib = SimpleNamespace(is_cuda=False)
    device = torch.device("cpu")

    def __init__(self):
        self.batches = []

    def nn_forward_planes(self, planes):
        self.batches.append(planes.shape[0])
        out = [op.fake_eval_from_planes(p.numpy()) for p in planes]
        return torch.as_tensor(np.stack([o[0] for o in out])), torch.as_tensor(np.array([o[1] for o in out], dtype=np.float32))

    def close(self):
        pass


def test_prediction_server_batches_and_scatters():
    from cczero_b200.api import CChessModelAPI
    api = CChessModelAPI(SimpleNamespace(model=None), agent_model=None, lib=SimpleNamespace(is_cuda=False), device="cpu")
    api.engine = StubEngine()
    pipes = [api.get_pipe(need_reload=False) for _ in range(3)]
    api.start(need_reload=False)
    states = [osenv.INIT_STATE, osenv.step(osenv.INIT_STATE, "1219"), osenv.step(osenv.INIT_STATE, "7747")]
    planes = [osenv.state_to_planes(s) for s in states]
    # three clients, requests of different sizes, two requests queued on one pipe
    pipes[0].send(planes[:2])
    pipes[1].send(planes[2:])
    pipes[2].send(planes)
    pipes[2].send(planes[:1])
    want = [op.fake_eval_from_planes(p) for p in planes]

    def recv(pipe):
        assert pipe.poll(30), api.last_error
        return pipe.recv()
    r0, r1, r2a, r2b = recv(pipes[0]), recv(pipes[1]), recv(pipes[2]), recv(pipes[2])
    for got, idx in ((r0, [0, 1]), (r1, [2]), (r2a, [0, 1, 2]), (r2b, [0])):
        assert len(got) == len(idx)
        for (p, v), i in zip(got, idx):
            assert isinstance(v, float) and p.dtype == np.float32 and (p == want[i][0]).all() and v == want[i][1]
    assert api.positions == 7 and sum(api.engine.batches) == 7 and len(api.engine.batches) <= 4
    # a client that goes away does not stop the server
    pipes[0].close()
    time.sleep(0.05)
    pipes[1].send(planes[:1])
    assert len(recv(pipes[1])) == 1
    api.close()
    assert not api.thread.is_alive() or api.done


def test_hot_reload_follows_the_digest(tmp_path):
    """api.py:76-88: when the best-model file changes (sha256 digest), the server loads it and re-uploads the weights."""
    from cczero_b200.api import CChessModelAPI
    from cczero_b200.model import CChessModel
    res = SimpleNamespace(model_best_config_path=str(tmp_path / "cfg.json"), model_best_weight_path=str(tmp_path / "w.npz"))
    cfg = SimpleNamespace(model=SimpleNamespace(cnn_filter_num=64, res_layer_num=1, value_fc_size=256, cnn_first_filter_size=5,
                                                cnn_filter_size=3, input_depth=14), resource=res)
    model = CChessModel(cfg).build(seed=1)
    model.save(res.model_best_config_path, res.model_best_weight_path)
    uploads = []
    eng = StubEngine()
    eng.set_weights = lambda w: uploads.append(float(w["value_out/kernel"].sum()))
    api = CChessModelAPI(cfg, model, lib=SimpleNamespace(is_cuda=False), device="cpu")
    api.engine = eng
    api.try_reload_model()
    assert uploads == []                                        # same digest: nothing to do
    other = CChessModel(cfg).build(seed=2)
    other.save(res.model_best_config_path, res.model_best_weight_path)
    api.try_reload_model()
    assert len(uploads) == 1 and model.digest == other.digest
    assert uploads[0] == float(other.weights["value_out/kernel"].sum())
    api.try_reload_model()
    assert len(uploads) == 1


def test_failed_reload_keeps_serving_and_retries(tmp_path):
    """A weight file the engine refuses (other geometry) or that is half written must leave the served model, its digest
    and its geometry untouched — the next check tries again instead of believing the reload happened."""
    from cczero_b200.api import CChessModelAPI
    from cczero_b200.model import CChessModel
    res = SimpleNamespace(model_best_config_path=str(tmp_path / "cfg.json"), model_best_weight_path=str(tmp_path / "w.npz"))
    mk = lambda f: SimpleNamespace(model=SimpleNamespace(cnn_filter_num=f, res_layer_num=1, value_fc_size=256, cnn_first_filter_size=5,
                                                         cnn_filter_size=3, input_depth=14), resource=res)
    cfg = mk(64)
    model = CChessModel(cfg).build(seed=1)
    model.save(res.model_best_config_path, res.model_best_weight_path)
    digest0 = model.digest
    uploads = []
    eng = StubEngine()
    eng.set_weights = lambda w: uploads.append(1)
    api = CChessModelAPI(cfg, model, lib=SimpleNamespace(is_cuda=False), device="cpu")
    api.engine = eng
    CChessModel(mk(128)).build(seed=2).save(res.model_best_config_path, res.model_best_weight_path)     # other geometry
    api.try_reload_model()
    assert uploads == [] and model.digest == digest0 and cfg.model.cnn_filter_num == 64
    assert model.weights["input_conv-5-64/kernel"].shape[-1] == 64
    with open(res.model_best_weight_path, "wb") as f:                                                   # half-written file
        f.write(b"PK\x03\x04 not a zip")
    api.try_reload_model()
    assert uploads == [] and model.digest == digest0
    good = CChessModel(mk(64)).build(seed=3)
    good.save(res.model_best_config_path, res.model_best_weight_path)
    api.try_reload_model()
    assert uploads == [1] and model.digest == good.digest
