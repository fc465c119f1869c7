#!/bin/bash
# round 2, GPU call R: final state — full GPU tier, smoke, bench (both arms), launch lists of the bench command, one ncu capture of
# the small-batch conv form
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $GOUT/r_pytest.log 2>&1
echo "pytest rc=$?" >> $GOUT/r_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $GOUT/r_smoke.log 2>&1
(time timeout 1200 python bench.py --steps 8 --warmup 4) > $GOUT/r_bench_c3.log 2>&1
(time timeout 1200 python bench.py --impl reference --steps 8 --warmup 4) > $GOUT/r_bench_ref.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $GOUT/r_launches_c3.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary --sims 64 > $GOUT/r_ncu_c3_list.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $GOUT/r_launches_c2.csv \
    python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-secondary > $GOUT/r_ncu_c2_list.log 2>&1
UCI_LOOPS=host timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_igemm3 -s 200 -c 2 -o $GOUT/r_conv_192_small \
    python tools/bench_uci.py > $GOUT/r_ncu_uci_full.log 2>&1
ls -la $GOUT
