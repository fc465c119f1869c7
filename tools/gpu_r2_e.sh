#!/bin/bash
# round 2, GPU call E: interleaved kernel-selection A/B, legacy heads parity, c3 bench A/B (same box, back to back)
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
(time timeout 600 python -m pytest tests/test_nn_gpu.py tests/test_keras_h5.py tests/test_adapters_gpu.py -m gpu -x -q -s) > $GOUT/e_pytest_nn.log 2>&1
echo "pytest rc=$?" >> $GOUT/e_pytest_nn.log
timeout 900 python tools/ab_nn.py 3 > $GOUT/e_ab_nn.log 2>&1
for v in 1 2; do
  CZ_EPI=2 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-secondary > $GOUT/e_bench_c3_epi2_$v.log 2>&1
  timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-secondary > $GOUT/e_bench_c3_auto_$v.log 2>&1
  CZ_EPI=3 CZ_NF=3 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-secondary > $GOUT/e_bench_c3_epi3nf3_$v.log 2>&1
done
ls -la $GOUT
