"""Play records and host staging.

Record format of the reference (worker/self_play.py:202-208 -> lib/data_helper.py:17-19): one JSON list per game,
`[init_state, [move, value], [move, -value], ...]` with `value` the result from red's view for the first entry and
alternating sign after it; moves are the 4-digit strings in the mover's own frame.
"""
import json
import os
from datetime import datetime, timedelta, timezone

import numpy as np
import torch

from .env import INIT_STATE, state_to_board
from .lib import BOARD_STRIDE, MAX_MOVES


def record_to_play_data(rec, init_state=INIT_STATE):
    """cz_drain_records entry -> the list self_play.py:202-208 builds."""
    data = [init_state]
    value = rec["value_red"]
    for m in rec["moves"]:
        data.append([m, value])
        value = -value
    return data


def write_play_data(play_data_dir, data, filename_tmpl="play_%s.json"):
    """save_play_data (self_play.py:214-227): Beijing-time stamped file, json.dump of the buffer."""
    os.makedirs(play_data_dir, exist_ok=True)
    bj = datetime.now(timezone.utc).astimezone(timezone(timedelta(hours=8)))
    path = os.path.join(play_data_dir, filename_tmpl % bj.strftime("%Y%m%d-%H%M%S.%f"))
    with open(path, "wt") as f:
        json.dump(data, f)
    return path


def init_boards_pinned(n_games):
    b = torch.from_numpy(np.tile(state_to_board(INIT_STATE), (n_games, 1)))
    return b.pin_memory() if torch.cuda.is_available() else b


class RootStage:
    """Pinned host buffers for the per-step host<->device traffic of the end-to-end path."""

    def __init__(self, engine):
        g = engine.n_games
        pin = engine.lib.is_cuda and torch.cuda.is_available()

        def mk(shape, dtype):
            t = torch.zeros(shape, dtype=dtype)
            return t.pin_memory() if pin else t
        self.boards = mk((g, BOARD_STRIDE), torch.uint8)
        self.n = mk((g, MAX_MOVES), torch.int32)
        self.moves = mk((g, MAX_MOVES), torch.int16)
        self.counts = mk((g,), torch.int32)
        self.sims = mk((g,), torch.int32)
        self.d2h_bytes_acc = 0

    @property
    def h2d_bytes(self):
        return self.boards.numel()

    @property
    def d2h_bytes(self):
        return self.boards.numel() + self.n.numel() * 4 + self.moves.numel() * 2 + self.counts.numel() * 4 + self.sims.numel() * 4


def decode_ring(ring_u8, count, layout):
    """A record ring as the collective delivered it (uint8 tensor / array) -> the dicts `Engine.drain_records` returns."""
    from .env import u16_to_move
    cap, stride, moves_off, _ = layout
    raw = ring_u8.cpu().numpy() if hasattr(ring_u8, "cpu") else np.asarray(ring_u8)
    count = min(int(count), cap)
    hdr = raw[:cap * 16].view(np.int32).reshape(cap, 4)
    moves = raw[moves_off:moves_off + cap * stride * 2].view(np.uint16).reshape(cap, stride)
    out = []
    for i in range(count):
        n_plies, value_red, game_index, flags = (int(x) for x in hdr[i])
        out.append({"n_plies": n_plies, "value_red": value_red, "game_index": game_index, "flags": flags,
                    "moves": [u16_to_move(int(v)) for v in moves[i, :n_plies]]})
    return out


def record_layout(engine):
    import ctypes as C
    a = np.zeros(4, dtype=np.int64)
    engine.lib.call("cz_record_layout", engine._h, C.c_void_p(a.ctypes.data))
    return tuple(int(x) for x in a)


def gather_records(engine, dist, world, decode_on=0, clear=True, warm=False):
    """all_gather (NCCL on GPUs, gloo in the CPU tests) of the finished-game record ring of every rank — SURVEY.md §8e:
    the only inter-GPU traffic of the path, the analogue of the reference uploading its play-data files
    (worker/self_play.py:228-241).  The ring lives inside the engine's workspace tensor, so the collective reads it in
    place; every rank must call this at the same point of its loop.  Returns (records, total): on rank `decode_on` the
    decoded records of ALL ranks as [(rank, record dict), ...] (None elsewhere: other ranks only forward), and the
    number of records gathered.  clear=True empties the local ring afterwards (its content now lives on rank decode_on).
    warm=True: also run the ring collective when no rank has a record yet (first call of a long run, bench warm-up)."""
    import ctypes as C
    ptr, nbytes, ready = C.c_void_p(0), C.c_uint64(0), C.c_int32(0)
    engine.lib.call("cz_record_buffer", engine._h, C.byref(ptr), C.byref(nbytes), C.byref(ready))
    off = ptr.value - engine.workspace.data_ptr()
    ring = engine.workspace[off:off + nbytes.value]
    count = torch.tensor([ready.value], device=engine.device, dtype=torch.int32)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count)
    counts = [int(c.item()) for c in counts]
    total = sum(counts)
    records = None
    if warm and total == 0:                     # warm-up call: run the ring collective once so that its one-off set-up (NCCL picks
        dist.all_gather([torch.empty_like(ring) for _ in range(world)], ring)   # channels per message size) is not paid later
    if total > 0:                               # same decision on every rank (they all hold the same counts)
        rings = [torch.empty_like(ring) for _ in range(world)]
        dist.all_gather(rings, ring)
        if dist.get_rank() == decode_on:
            layout = record_layout(engine)
            records = [(r, rec) for r in range(world) for rec in decode_ring(rings[r], counts[r], layout)]
    elif dist.get_rank() == decode_on:
        records = []
    if clear:
        engine.lib.call("cz_clear_records", engine._h)
    return records, total


# ---- trainer-side view of the records (worker/optimize.py:223-292, lib/data_helper.py:11-24) -------------------
def get_game_data_filenames(rc):
    from glob import glob
    return sorted(glob(os.path.join(rc.play_data_dir, rc.play_data_filename_tmpl % "*")))


def read_game_data_from_file(path):
    with open(path, "rt") as f:
        return json.load(f)


def expanding_data(data, env, use_history=False):
    """expanding_data + convert_to_trainging_data (optimize.py:234-281): one play record
    `[init_state, [move, value], ...]` -> (planes f32 [T,14,10,9], one-hot policy f32 [T,2086], value f32 [T]).
    use_history: planes f32 [T,28,10,9], planes 14-27 of sample i = the position of sample i-2 (history[0:2i+1][-5],
    optimize.py:264-267), zero for the first two.  The positions are replayed and encoded by the rules kernels
    (`env` is a StaticEnv)."""
    from .env import move_to_u16
    moves = [item[0] for item in data[1:]]
    values = np.asarray([item[1] for item in data[1:]], dtype=np.float32)
    t = len(moves)
    boards = env.boards_from_states([data[0]])
    seq = [boards]
    for m in moves[:-1]:
        boards, _ = env.step_batch(boards, env.moves_tensor([m]))
        seq.append(boards)
    if t == 0:
        return (np.zeros((0, 28 if use_history else 14, 10, 9), np.float32), np.zeros((0, len(env.labels)), np.float32), values)
    planes = env.planes_batch(torch.cat(seq, dim=0)).cpu().numpy()
    if use_history:
        hist = np.zeros_like(planes)
        hist[2:] = planes[:-2]
        planes = np.concatenate([planes, hist], axis=1)
    policy = np.zeros((t, len(env.labels)), dtype=np.float32)
    lut = env.label_lut
    for i, m in enumerate(moves):
        v = move_to_u16(m)
        lab = int(lut[(v >> 8) * 90 + (v & 0xFF)])
        if lab < 0:
            raise ValueError(f"move {m} is not an action label")
        policy[i, lab] = 1
    return planes, policy, values


def flip_policy(pol, env):
    """lookup_tables.py:134-141: re-index a 2086-vector from black's move labels to red's
    (out[i] = pol[index of flip_move(label_i)])."""
    from .env import flip_move
    if not hasattr(env, "_unflipped_index"):
        lookup = {m: i for i, m in enumerate(env.labels)}
        env._unflipped_index = np.asarray([lookup[flip_move(m)] for m in env.labels])
    return np.asarray(pol)[env._unflipped_index]


def build_policy(action, flip, env):
    """optimize.py:283-292 / self_play.py:253-262: one-hot over the action labels, optionally seen from the other side."""
    policy = np.zeros(len(env.labels))
    policy[env.labels.index(action)] = 1
    if flip:
        policy = flip_policy(policy, env)
    return list(policy)
