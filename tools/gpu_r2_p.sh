#!/bin/bash
# round 2, GPU call P: root noise drawn ahead by K warps per game (k_noise_fill) + per-search noise streams; PDL on by default
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $GOUT/p_pytest.log 2>&1
echo "pytest rc=$?" >> $GOUT/p_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $GOUT/p_smoke.log 2>&1
for v in 0 1 0 1; do
  CZ_NOISE_AHEAD=$v timeout 300 python tools/bench_tree.py > $GOUT/p_bench_tree_ahead$v.log 2>&1
  CZ_NOISE_AHEAD=$v UCI_LOOPS=while timeout 300 python tools/bench_uci.py > $GOUT/p_bench_uci_ahead$v.log 2>&1
  CZ_NOISE_AHEAD=$v timeout 300 python bench.py --workload c2 --steps 30 --warmup 5 --no-cpu --no-secondary > $GOUT/p_bench_c2_ahead$v.log 2>&1
  for f in tree uci c2; do echo "ahead=$v $(tail -1 $GOUT/p_bench_${f}_ahead$v.log)" >> $GOUT/p_ab_$f.log; done
done
ls -la $GOUT
