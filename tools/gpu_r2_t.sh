#!/bin/bash
# round 2, GPU call T: L2 evict_first policy on the fp32 skip stream of conv2 (igemm3 form): parity, then interleaved A/B
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
(time CZ_EPI=3 CZ_L2HINT=1 timeout 600 python -m pytest tests/test_nn_gpu.py -m gpu -x -q) > $GOUT/t_pytest_hint.log 2>&1
echo "pytest rc=$?" >> $GOUT/t_pytest_hint.log
AB_SHAPES=c3 AB_ONLY="skip default,hint" timeout 900 python tools/ab_nn.py 3 > $GOUT/t_ab_nn.log 2>&1
for v in "" "CZ_EPI=3 CZ_L2HINT=1" "" "CZ_EPI=3 CZ_L2HINT=1"; do
  env $v timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-secondary > $GOUT/t_bench_c3.log 2>&1
  echo "[$v] $(tail -1 $GOUT/t_bench_c3.log)" >> $GOUT/t_ab_c3.log
done
ls -la $GOUT
