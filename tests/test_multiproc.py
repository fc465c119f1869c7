"""N > 1 path on CPU: two ranks over gloo, each with its own engine (CPU SIMT-emulation build), disjoint RNG streams,
records gathered with the same `records.gather_records` the NCCL path uses."""
import os
import socket
import sys

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
        if finished >= 2:
            break
    total = records.gather_records(eng, dist, world)
    mine = torch.tensor([finished], dtype=torch.int64)
    dist.all_reduce(mine)
    out.put((rank, finished, total, int(mine.item()), first_moves))
    eng.close()
    dist.destroy_process_group()


def test_two_ranks_gloo(emul_lib):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, emul_lib.path, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, f0, t0, s0, m0), (r1, f1, t1, s1, m1) = res
    assert t0 == t1 == s0 == s1 == f0 + f1          # every rank sees all records after the gather
    assert f0 >= 2 and f1 >= 2
    assert m0 != m1                                  # rank-specific Philox sub-streams: different root noise
