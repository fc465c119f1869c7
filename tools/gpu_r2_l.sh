#!/bin/bash
# round 2, GPU call L: noinline rules functions: parity, tree microbench, c2, k_apply_wave profile
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
(time timeout 900 python -m pytest tests/test_env.py tests/test_search.py tests/test_selfplay.py tests/test_arena.py tests/test_games_golden.py tests/test_compact.py tests/test_uci.py tests/test_adapters_gpu.py -m gpu -x -q) > $GOUT/l_pytest.log 2>&1
echo "pytest rc=$?" >> $GOUT/l_pytest.log
timeout 300 python tools/bench_tree.py > $GOUT/l_bench_tree.log 2>&1
timeout 300 python tools/bench_tree.py 256 200 8 > $GOUT/l_bench_tree_256.log 2>&1
for v in 1 2; do timeout 300 python bench.py --workload c2 --steps 30 --warmup 5 --no-cpu --no-secondary > $GOUT/l_bench_c2_$v.log 2>&1; done
timeout 600 python tools/bench_uci.py > $GOUT/l_bench_uci.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_apply_wave -s 12 -c 1 -o $GOUT/l_apply_wave \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary --sims 128 > $GOUT/l_ncu_wave.log 2>&1
ls -la $GOUT
