#!/bin/bash
# round 2, GPU call M: where a single-game (UCI) wave spends its 700 us: launch list of the host-loop form
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
UCI_LOOPS=host timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file $GOUT/m_launches_uci.csv \
    python tools/bench_uci.py > $GOUT/m_uci_ncu.log 2>&1
timeout 300 python tools/bench_uci.py 192 10 8 40 > $GOUT/m_bench_uci_k40.log 2>&1
ls -la $GOUT
