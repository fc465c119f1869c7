#!/bin/bash
# round 2, GPU call O: programmatic dependent launch of the residual convs (CZ_PDL=1) — correctness in every loop form, then A/B
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
(time CZ_PDL=1 timeout 900 python -m pytest tests/test_nn_gpu.py tests/test_adapters_gpu.py tests/test_uci.py -m gpu -x -q) > $GOUT/o_pytest_pdl.log 2>&1
echo "pytest rc=$?" >> $GOUT/o_pytest_pdl.log
for v in 0 1 0 1; do
  CZ_PDL=$v UCI_LOOPS=while,graph timeout 300 python tools/bench_uci.py > $GOUT/o_bench_uci_pdl$v.log 2>&1
  CZ_PDL=$v timeout 300 python bench.py --workload c2 --steps 30 --warmup 5 --no-cpu --no-secondary > $GOUT/o_bench_c2_pdl$v.log 2>&1
  for f in uci c2; do tail -1 $GOUT/o_bench_${f}_pdl$v.log >> $GOUT/o_ab_$f.log; done
done
for v in 0 1; do CZ_PDL=$v timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-secondary > $GOUT/o_bench_c3_pdl$v.log 2>&1; done
timeout 300 python tools/bench_tree.py 2048 800 8 > $GOUT/o_bench_tree_2048.log 2>&1
timeout 300 python tools/bench_tree.py 4096 800 8 > $GOUT/o_bench_tree_4096.log 2>&1
ls -la $GOUT
