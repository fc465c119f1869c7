#!/bin/bash
# round 2, GPU call S: wide skip stream as fp16 + 8 (CZ_SKIP_FORMAT=split8): parity, then interleaved A/B against fp32 and fp16-only
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
(time CZ_SKIP_FORMAT=split8 timeout 600 python -m pytest tests/test_nn_gpu.py tests/test_keras_h5.py -m gpu -x -q) > $GOUT/s_pytest_split8.log 2>&1
echo "pytest rc=$?" >> $GOUT/s_pytest_split8.log
CZ_SKIP_FORMAT=split8 timeout 300 python tools/nn_error_report.py > $GOUT/s_nn_error_split8.log 2>&1
CZ_SKIP_FORMAT=fp32 timeout 300 python tools/nn_error_report.py > $GOUT/s_nn_error_fp32.log 2>&1
AB_ONLY="skip" timeout 900 python tools/ab_nn.py 3 > $GOUT/s_ab_nn.log 2>&1
for v in fp32 split8 fp32 split8; do
  CZ_SKIP_FORMAT=$v timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-secondary > $GOUT/s_bench_c3_$v.log 2>&1
  echo "$v $(tail -1 $GOUT/s_bench_c3_$v.log)" >> $GOUT/s_ab_c3.log
done
ls -la $GOUT
