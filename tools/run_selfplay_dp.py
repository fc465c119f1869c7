"""Data-parallel self-play through the drop-in entry point (cczero_b200.self_play.start, the role of the reference's
`run.py self` -> worker/self_play.start): one process per GPU under torchrun, finished-game rings all_gathered over NCCL,
rank 0 writes the reference-layout play-data files.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/run_selfplay_dp.py [out_dir] [games]
"""
import glob
import json
import os
import sys
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "/tmp/cz_dp_selfplay"
    games = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    from cczero_b200 import self_play
    play = SimpleNamespace(max_processes=1, simulation_num_per_move=64, search_threads=8, virtual_loss=3, c_puct=1.5, noise_eps=0.25,
                           dirichlet_alpha=0.2, tau_decay_rate=0.98, resign_threshold=-0.92, enable_resign_rate=0.1, min_resign_turn=20,
                           max_game_length=30)
    cfg = SimpleNamespace(play=play, model=SimpleNamespace(cnn_filter_num=128, res_layer_num=4, value_fc_size=256, input_depth=14,
                                                           cnn_first_filter_size=5, cnn_filter_size=3),
                          play_data=SimpleNamespace(nb_game_in_file=1),
                          resource=SimpleNamespace(play_data_dir=os.path.join(out, "play_data"), play_data_filename_tmpl="play_%s.json",
                                                   model_best_config_path=os.path.join(out, "model", "model_best_config.json"),
                                                   model_best_weight_path=os.path.join(out, "model", "model_best_weight.npz")))
    t0 = time.time()
    stored = self_play.start(cfg, games_per_process=256, max_games=games, flush_plies=8)
    dt = time.time() - t0
    if rank == 0:
        files = glob.glob(os.path.join(cfg.resource.play_data_dir, "play_*.json"))
        ok = all(isinstance(json.load(open(f))[0], str) for f in files[:20])
        print(json.dumps({"world": world, "games_stored": stored, "files": len(files), "seconds": round(dt, 1), "files_parse": ok}))


if __name__ == "__main__":
    main()
