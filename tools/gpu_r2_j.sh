#!/bin/bash
# round 2, GPU call J: two M-tiles per CTA at C <= 128 (k_igemm3<N, 2>): parity, interleaved A/B, c2 bench, ncu
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
(time timeout 900 python -m pytest tests/test_igemm_gpu.py tests/test_nn_gpu.py tests/test_adapters_gpu.py tests/test_full_size_gpu.py -m gpu -x -q) > $GOUT/j_pytest.log 2>&1
echo "pytest rc=$?" >> $GOUT/j_pytest.log
AB_ONLY="auto" AB_SHAPES=small timeout 600 python tools/ab_nn.py 3 > $GOUT/j_ab_nn.log 2>&1
for v in 1 2; do
  CZ_MT=1 timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu --no-secondary > $GOUT/j_bench_c2_mt1_$v.log 2>&1
  timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu --no-secondary > $GOUT/j_bench_c2_mt2_$v.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_igemm3 -s 46 -c 2 -o $GOUT/j_conv_128_mt2 \
    python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-secondary > $GOUT/j_ncu_c2_full.log 2>&1
ls -la $GOUT
