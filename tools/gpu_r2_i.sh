#!/bin/bash
# round 2, GPU call I: final state — full GPU tier, smoke, bench (both arms), launch lists + ncu captures of the shipped kernels
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
(time timeout 1500 python -m pytest tests -m gpu -x -q -s) > $GOUT/i_pytest.log 2>&1
echo "pytest rc=$?" >> $GOUT/i_pytest.log
timeout 300 python __graft_entry__.py smoke > $GOUT/i_smoke.log 2>&1
(time timeout 900 python bench.py --steps 8 --warmup 4) > $GOUT/i_bench_c3.log 2>&1
(time timeout 600 python bench.py --impl reference --steps 6 --warmup 1) > $GOUT/i_bench_ref.log 2>&1
timeout 600 python tools/bench_uci.py > $GOUT/i_bench_uci.log 2>&1
timeout 300 python tools/bench_tree.py > $GOUT/i_bench_tree.log 2>&1
# launch lists (shares) of the bench command, short
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $GOUT/i_launches_c3.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary --sims 64 > $GOUT/i_ncu_c3_list.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $GOUT/i_launches_c2.csv \
    python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-secondary > $GOUT/i_ncu_c2_list.log 2>&1
# full captures: conv1 (k_igemm3) + conv2 (k_igemm2) of one residual block at c3, the same pair at c2, the fused tree kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_igemm[23] -s 46 -c 2 -o $GOUT/i_conv_256 \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary --sims 64 > $GOUT/i_ncu_c3_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_igemm[23] -s 46 -c 2 -o $GOUT/i_conv_128 \
    python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-secondary > $GOUT/i_ncu_c2_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_apply_wave -s 12 -c 1 -o $GOUT/i_apply_wave \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary --sims 128 > $GOUT/i_ncu_wave.log 2>&1
ls -la $GOUT
