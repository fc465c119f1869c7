#!/bin/bash
# round 2, GPU call D: 16-column TMA epilogue, WHILE-graph loop, noise-table fix: full parity tier + A/B microbench + bench
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
(time timeout 1500 python -m pytest tests -m gpu -x -q -s) > $GOUT/d_pytest.log 2>&1
echo "pytest rc=$?" >> $GOUT/d_pytest.log
timeout 300 python __graft_entry__.py smoke > $GOUT/d_smoke.log 2>&1
CZ_EPI=3 timeout 300 python tools/bench_nn.py > $GOUT/d_bench_nn_epi3.log 2>&1
CZ_EPI=3 CZ_NF=3 timeout 300 python tools/bench_nn.py > $GOUT/d_bench_nn_epi3_nf3.log 2>&1
timeout 300 python tools/bench_nn.py > $GOUT/d_bench_nn_auto.log 2>&1
(time timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu) > $GOUT/d_bench_c3.log 2>&1
CZ_EPI=3 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-secondary > $GOUT/d_bench_c3_epi3.log 2>&1
CZ_SEARCH_LOOP=graph timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-secondary > $GOUT/d_bench_c3_subgraphs.log 2>&1
timeout 300 python bench.py --workload c2 --steps 12 --warmup 4 --no-cpu --no-secondary > $GOUT/d_bench_c2.log 2>&1
timeout 600 python tools/bench_uci.py > $GOUT/d_bench_uci.log 2>&1
timeout 300 python tools/bench_tree.py > $GOUT/d_bench_tree.log 2>&1
CZ_EPI=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_igemm3 -s 45 -c 1 -o $GOUT/d_igemm3_256_conv2 \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary --sims 64 > $GOUT/d_ncu_c3_full.log 2>&1
ls -la $GOUT
