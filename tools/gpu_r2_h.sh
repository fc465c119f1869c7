#!/bin/bash
# round 2, GPU call H: bitboard movegen on the device (parity incl. the 1e5 sweep), tree microbench / c2 after, k_wave source profile
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
(time timeout 900 python -m pytest tests/test_env.py tests/test_search.py tests/test_selfplay.py tests/test_arena.py tests/test_games_golden.py tests/test_compact.py tests/test_uci.py -m gpu -x -q -s) > $GOUT/h_pytest.log 2>&1
echo "pytest rc=$?" >> $GOUT/h_pytest.log
timeout 300 python tools/bench_tree.py > $GOUT/h_bench_tree.log 2>&1
timeout 300 python tools/bench_tree.py 256 200 8 > $GOUT/h_bench_tree_256.log 2>&1
timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu --no-secondary > $GOUT/h_bench_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wave -s 60 -c 1 -o $GOUT/h_kwave python tools/bench_tree.py > $GOUT/h_ncu_kwave.log 2>&1
ls -la $GOUT
