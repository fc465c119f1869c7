#!/bin/bash
# round 2, GPU call N: small-batch network path (64-column conv tiles, sliced first conv, block-per-position heads)
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
(time timeout 900 python -m pytest tests/test_nn_gpu.py tests/test_adapters_gpu.py tests/test_uci.py tests/test_search.py -m gpu -x -q) > $GOUT/n_pytest.log 2>&1
echo "pytest rc=$?" >> $GOUT/n_pytest.log
timeout 600 python tools/bench_uci.py > $GOUT/n_bench_uci.log 2>&1
CZ_NSPLIT=0 UCI_LOOPS=while timeout 300 python tools/bench_uci.py > $GOUT/n_bench_uci_nosplit.log 2>&1
UCI_LOOPS=host timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 600 --csv --log-file $GOUT/n_launches_uci.csv \
    python tools/bench_uci.py > $GOUT/n_uci_ncu.log 2>&1
timeout 300 python bench.py --workload c2 --steps 30 --warmup 5 --no-cpu --no-secondary > $GOUT/n_bench_c2.log 2>&1
ls -la $GOUT
