#!/bin/bash
# round 2, GPU call X: last sanity of the final commit (GPU tier, smoke, a short default bench)
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
(time timeout 1200 python -m pytest tests -m gpu -x -q) > $GOUT/x_pytest.log 2>&1
echo "pytest rc=$?" >> $GOUT/x_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $GOUT/x_smoke.log 2>&1
(time timeout 900 python bench.py --steps 3 --warmup 3 --cpu-seconds 10) > $GOUT/x_bench_c3.log 2>&1
ls -la $GOUT
