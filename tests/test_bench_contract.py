"""bench.py JSON contract (CPU side): the reference arm runs without a GPU and prints one line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    env = dict(os.environ, CZ_BENCH_CPU_PROCS="2", CZ_BENCH_CPU_NN_THREADS="2", CZ_BENCH_CPU_WINDOW="3", CZ_BENCH_CPU_SETTLE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "2",
                          "--warmup", "0", "--no-secondary"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "mcts_sims_per_sec" and line["unit"] == "sims/s"
    from oracle import ref_selfplay_bench as rb
    cb = line["cpu_baseline"]
    if rb.available():        # oracle/_ref built (build() does it wherever /root/reference exists): the reference's own plumbing
        assert cb["kind"] == "reference" and cb["cores"] == 4 and cb["max_processes"] == 2 and cb["nn_threads"] == 2
        assert "SelfPlayWorker" in cb["sample"] and len(cb["window_rates"]) == 2
    else:                     # a checkout that never saw the reference: the oracle port, said so
        assert cb["kind"] == "port"
    assert line["value"] > 0 and line["value"] == cb["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_ours_arm_refuses_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True,
                         text=True, timeout=200, cwd=ROOT)
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)
