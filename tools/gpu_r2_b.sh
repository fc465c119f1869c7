#!/bin/bash
# round 2, GPU call B: device-driven search loop + legal priors from logits: parity tests, bench c3 / c2, UCI latency
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
(time timeout 1500 python -m pytest tests -m gpu -x -q -s) > $GOUT/b_pytest.log 2>&1
echo "pytest rc=$?" >> $GOUT/b_pytest.log
timeout 300 python __graft_entry__.py smoke > $GOUT/b_smoke.log 2>&1
(time timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu) > $GOUT/b_bench_c3.log 2>&1
CZ_SEARCH_LOOP=host timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-secondary > $GOUT/b_bench_c3_hostloop.log 2>&1
timeout 300 python bench.py --workload c2 --steps 12 --warmup 4 --no-cpu --no-secondary > $GOUT/b_bench_c2.log 2>&1
CZ_SEARCH_LOOP=host timeout 300 python bench.py --workload c2 --steps 12 --warmup 4 --no-cpu --no-secondary > $GOUT/b_bench_c2_hostloop.log 2>&1
timeout 600 python tools/bench_uci.py > $GOUT/b_bench_uci.log 2>&1
timeout 600 python tools/bench_uci.py 256 20 8 10 > $GOUT/b_bench_uci_256x20.log 2>&1
ls -la $GOUT
