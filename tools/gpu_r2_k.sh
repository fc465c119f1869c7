#!/bin/bash
# round 2, GPU call K: final tree — full GPU tier, smoke, bench with secondary, c2 launch list
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
(time timeout 1500 python -m pytest tests -m gpu -x -q -s) > $GOUT/k_pytest.log 2>&1
echo "pytest rc=$?" >> $GOUT/k_pytest.log
timeout 300 python __graft_entry__.py smoke > $GOUT/k_smoke.log 2>&1
(time timeout 900 python bench.py --steps 8 --warmup 4) > $GOUT/k_bench_c3.log 2>&1
timeout 300 python bench.py --workload c2 --steps 30 --warmup 5 --no-cpu --no-secondary > $GOUT/k_bench_c2.log 2>&1
timeout 600 python tools/bench_uci.py > $GOUT/k_bench_uci.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $GOUT/k_launches_c2.csv \
    python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-secondary > $GOUT/k_ncu_c2_list.log 2>&1
ls -la $GOUT
