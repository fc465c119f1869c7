"""CPU arm of bench.py: the reference's OWN `run.py self` plumbing, timed.  TEST / BENCH INFRASTRUCTURE.

What runs (all of it the reference's byte-compiled code from oracle/_ref, see oracle/build_ref.py):

    worker/self_play.start(config)                              (manager.py:72-80 calls exactly this)
      -> load_model -> CChessModel.get_pipes -> CChessModelAPI.predict_batch_worker   (agent/api.py:37-74, one thread)
      -> ProcessPoolExecutor(max_processes) x SelfPlayWorker.start -> start_game      (worker/self_play.py:48-212)
           -> CChessPlayer.action / MCTS_search / update_tree                        (agent/player.py:145-373)
           <-> multiprocessing.Pipe <-> the prediction thread

Substitutions, because TensorFlow 1.3 / Keras 2.0.8 cannot be installed here — and nothing else:
  * `tensorflow` / `keras.*` imports are satisfied by empty stand-in modules;
  * `CChessModel.build` sets `.model` to the fp32 PyTorch-CPU restatement of agent/model.py:32-83 (oracle/model.py
    TorchNet, Keras-equivalent random init) behind Keras' `predict_on_batch(np.float32[B,14,10,9]) -> (policy, value)`;
    `.save` is a no-op (there is no Keras model to serialise);
  * `config.model.input_depth = 28`: the reference's load_model returns use_history=True for a freshly built network
    (self_play.py:40-43) while its configs say input_depth = 14 (configs/mini.py:82) — as written `run.py self --new` feeds
    28-plane batches to a 14-plane Keras input.  The stand-in network is therefore built for the 28 planes the players
    actually send (+0.2 % FLOPs on 256x20); the config value is set, no reference code is changed;
  * the start-up stagger `sleep((pid % ran) * 10)` (self_play.py:74-75) is skipped;
  * counting only: `CChessPlayer.update_tree` is wrapped to count finished simulations (one call = one decrement of
    num_task, player.py:369-371) and `predict_on_batch` counts evaluated positions, both into shared counters.
The whole thing runs in a child process group that is killed when the measurement is over (the reference's worker loop
never returns).  `free_nn=True` swaps the network for a constant-output stub: the tree-code ceiling.
"""
import contextlib
import multiprocessing as mp
import os
import signal
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_BUILD = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF_BUILD, "cchess_alphazero", "worker", "self_play.pyc"))


class _Graph:
    def as_default(self):
        return contextlib.nullcontext()


def _child(conf, sims_done, positions, batches, ready):
    os.setsid()                                            # own process group: the parent kills the whole tree
    os.environ["OMP_NUM_THREADS"] = str(conf["nn_threads"])
    os.environ["MKL_NUM_THREADS"] = str(conf["nn_threads"])
    scratch = tempfile.mkdtemp(prefix="cz_refbench_")
    os.environ["PROJECT_DIR"] = scratch
    os.environ["DATA_DIR"] = os.path.join(scratch, "data")
    os.environ.pop("MODEL_DIR", None)
    sys.dont_write_bytecode = True
    for p in (os.path.join(REF_BUILD, "cchess_alphazero"), REF_BUILD, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import logging
    logging.disable(logging.CRITICAL)
    import numpy as np
    import torch
    torch.set_num_threads(conf["nn_threads"])
    from oracle.ref_worker_harness import install_shims
    install_shims()                                        # tensorflow / keras -> empty stand-ins
    from oracle import model as om
    import cchess_alphazero.agent.model as ref_model
    import cchess_alphazero.agent.player as ref_player
    import cchess_alphazero.worker.self_play as ref_self_play
    from cchess_alphazero.config import Config

    config = Config(conf["config_type"])
    config.resource.create_directories()
    pc, mc = config.play, config.model
    pc.max_processes = conf["procs"]
    for k, v in conf["play"].items():
        setattr(pc, k, v)
    for k, v in conf["model"].items():
        setattr(mc, k, v)
    mc.input_depth = 28                                    # what a fresh model's players send (use_history=True, see header)
    config.internet.distributed = False
    config.opts.device_list = "0"
    config.opts.new = True

    class _Net:
        def __init__(self):
            if conf["free_nn"]:
                self.p = np.full((1, 2086), 1.0 / 2086, dtype=np.float32)
                self.net = None
            else:
                self.net = om.TorchNet(om.init_weights(mc.cnn_filter_num, mc.res_layer_num, mc.value_fc_size, seed=0,
                                                       in_planes=mc.input_depth), mc.res_layer_num)

        def predict_on_batch(self, data):
            n = len(data)
            with positions.get_lock():
                positions.value += n
                batches.value += 1
            if self.net is None:
                return np.repeat(self.p, n, axis=0), np.zeros((n, 1), dtype=np.float32)
            return self.net.predict_on_batch(np.asarray(data, dtype=np.float32))

    def build(self):
        self.model = _Net()
        self.graph = _Graph()
    ref_model.CChessModel.build = build
    ref_model.CChessModel.save = lambda self, config_path, weight_path: None

    orig_update = ref_player.CChessPlayer.update_tree

    def counted_update(self, p, v, history):
        orig_update(self, p, v, history)
        with sims_done.get_lock():
            sims_done.value += 1
    ref_player.CChessPlayer.update_tree = counted_update

    real_sleep = ref_self_play.sleep
    ref_self_play.sleep = lambda s: real_sleep(s) if s < 1 else None      # only the (pid % ran) * 10 start-up stagger is dropped
    ready.value = 1
    ref_self_play.start(config)                            # never returns (SelfPlayWorker.start loops for ever)


class ReferenceSelfPlay:
    """Persistent run of the reference's self-play; `window(seconds)` -> (simulations, positions, batches, seconds)."""

    def __init__(self, config_type="mini", procs=1, nn_threads=1, play=None, model=None, free_nn=False):
        if not available():
            raise RuntimeError("oracle/_ref is not built (python -m oracle.build_ref needs /root/reference)")
        self.conf = {"config_type": config_type, "procs": int(procs), "nn_threads": int(nn_threads), "play": play or {},
                     "model": model or {}, "free_nn": bool(free_nn)}
        ctx = mp.get_context("fork")
        self.sims, self.pos, self.bat, self.ready = ctx.Value("q", 0), ctx.Value("q", 0), ctx.Value("q", 0), ctx.Value("i", 0)
        self.proc = ctx.Process(target=_child, args=(self.conf, self.sims, self.pos, self.bat, self.ready), daemon=False)
        self.proc.start()

    def wait_started(self, timeout=180.0, min_sims=1):
        t0 = time.time()
        while time.time() - t0 < timeout:
            if not self.proc.is_alive():
                raise RuntimeError("reference self-play process died during start-up")
            if self.sims.value >= min_sims:
                return time.time() - t0
            time.sleep(0.05)
        raise RuntimeError("reference self-play produced no simulation within %.0f s" % timeout)

    def snapshot(self):
        return self.sims.value, self.pos.value, self.bat.value, time.time()

    def window(self, seconds):
        s0, p0, b0, t0 = self.snapshot()
        time.sleep(seconds)
        s1, p1, b1, t1 = self.snapshot()
        if not self.proc.is_alive():
            raise RuntimeError("reference self-play process died")
        return s1 - s0, p1 - p0, b1 - b0, t1 - t0

    def close(self):
        if self.proc.is_alive():
            try:
                os.killpg(self.proc.pid, signal.SIGKILL)  # exactly the process group this object started
            except ProcessLookupError:
                pass
        self.proc.join(timeout=10)


def describe(conf_play, conf_model, procs, nn_threads, free_nn=False):
    net = "constant-output stub (free NN)" if free_nn else (f"{conf_model['cnn_filter_num']}x{conf_model['res_layer_num']} fp32 torch-CPU "
                                                           f"predict_on_batch, {nn_threads} intra-op threads")
    return (f"unmodified reference plumbing (self_play.start -> SelfPlayWorker/CChessPlayer <-> Pipe <-> CChessModelAPI thread), "
            f"max_processes={procs}, search_threads={conf_play['search_threads']}, {conf_play['simulation_num_per_move']} sims/move, {net}")


if __name__ == "__main__":
    import json
    procs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    thr = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    secs = float(sys.argv[3]) if len(sys.argv) > 3 else 10
    r = ReferenceSelfPlay("mini", procs, thr, play={"simulation_num_per_move": 100, "search_threads": 10},
                          model={"cnn_filter_num": 256, "res_layer_num": 7})
    try:
        print("start-up", r.wait_started())
        s, p, b, dt = r.window(secs)
        print(json.dumps({"sims_per_s": s / dt, "positions_per_s": p / dt, "mean_batch": p / max(1, b), "seconds": dt}))
    finally:
        r.close()
