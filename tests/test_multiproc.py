"""N > 1 path on CPU: two ranks over gloo, each with its own engine (CPU SIMT-emulation build), disjoint RNG streams,
records gathered with the same `records.gather_records` the NCCL path uses, and the data-parallel `self_play.start`
launch (rank 0 decodes every rank's ring and writes the play-data files)."""
import glob
import json
import os
import socket
import sys
from types import SimpleNamespace

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, emul_path, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cczero_b200.engine import Engine
    from cczero_b200.lib import CzLib
    from cczero_b200 import records
    from tests.search_checks import eval_planes
    lib = CzLib(emul_path)
    eng = Engine(lib, "cpu", n_games=2, sims_per_move=12, leaves_per_round=4, noise_mode=1, noise_eps=0.25,
                 max_game_length=8, seed=3, rank=rank, max_nodes_per_game=2048)
    eng.reset()
    first_moves = []
    finished = 0
    for ply in range(40):
        eng.search_external(eval_planes, None)
        if ply == 0:
            r = eng.root(0)
            first_moves = list(zip(r["moves"], r["n"]))
        finished += eng.play_move()
        if finished >= 2 + rank:                     # ranks hold different numbers of records
            break
    gathered, total = records.gather_records(eng, dist, world, clear=False)
    mine = eng.drain_records()                       # what this rank's ring held
    again, total2 = records.gather_records(eng, dist, world, warm=True)   # rings are empty now: nothing is shipped twice (warm: the ring collective still runs once)
    out.put((rank, finished, total, total2, mine, gathered, first_moves))
    eng.close()
    dist.destroy_process_group()


def _spawn(target, args_of_rank, world=2, timeout=300):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, world, port) + args_of_rank + (q,)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=timeout) for _ in procs), key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_two_ranks_gloo(emul_lib):
    (r0, f0, t0, z0, mine0, gath0, m0), (r1, f1, t1, z1, mine1, gath1, m1) = _spawn(_worker, (emul_lib.path,))
    assert t0 == t1 == f0 + f1 and z0 == z1 == 0    # every rank learns the total; a second gather finds empty rings
    assert f0 >= 2 and f1 >= 3 and len(mine0) == f0 and len(mine1) == f1
    assert gath1 is None                             # only rank 0 decodes
    # CONTENTS: what rank 0 decoded from the collective == what each rank drained from its own ring
    assert [rec for r, rec in gath0 if r == 0] == mine0
    assert [rec for r, rec in gath0 if r == 1] == mine1
    assert all(rec["n_plies"] == len(rec["moves"]) > 0 for _, rec in gath0)
    assert m0 != m1                                  # rank-specific Philox sub-streams: different root noise


def _start_worker(rank, world, port, emul_path, data_dir, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from cczero_b200 import self_play
    from cczero_b200.lib import CzLib
    from tests.search_checks import eval_planes
    play = SimpleNamespace(max_processes=1, simulation_num_per_move=8, search_threads=4, virtual_loss=3, c_puct=1.5, noise_eps=0.25,
                           dirichlet_alpha=0.2, tau_decay_rate=0.9, resign_threshold=-0.98, enable_resign_rate=0.0,
                           min_resign_turn=40, max_game_length=4)
    cfg = SimpleNamespace(play=play, model=SimpleNamespace(cnn_filter_num=64, res_layer_num=1, value_fc_size=256, input_depth=14,
                                                           cnn_first_filter_size=5, cnn_filter_size=3),
                          play_data=SimpleNamespace(nb_game_in_file=1),
                          resource=SimpleNamespace(play_data_dir=data_dir, play_data_filename_tmpl="play_%s.json",
                                                   model_best_config_path=os.path.join(data_dir, "m", "cfg.json"),
                                                   model_best_weight_path=os.path.join(data_dir, "m", "w.npz")))
    stored = self_play.start(cfg, games_per_process=3, max_games=8, flush_plies=2, lib=CzLib(emul_path), device="cpu",
                             evaluate_planes=eval_planes)
    out.put((rank, stored))


def test_data_parallel_self_play_start(emul_lib, tmp_path):
    """`self_play.start` under a 2-rank launch: both ranks play, rank 0 alone writes reference-layout files for the games
    of BOTH ranks, and the ranks stop together."""
    res = _spawn(_start_worker, (emul_lib.path, str(tmp_path)))
    assert res[0][1] == res[1][1] >= 8
    files = sorted(glob.glob(str(tmp_path / "play_*.json")))
    assert len(files) == res[0][1]
    for f in files:
        data = json.load(open(f))
        assert isinstance(data[0], str) and len(data) >= 2 and all(len(m) == 4 and v in (-1, 0, 1) for m, v in data[1:])
        assert [v for _, v in data[1:]] == [data[1][1] * (-1) ** i for i in range(len(data) - 1)]   # alternating sign
