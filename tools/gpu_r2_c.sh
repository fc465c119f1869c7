#!/bin/bash
# round 2, GPU call C: all-TMA conv epilogue (k_igemm3) + new heads / first-conv kernels: parity, microbench A/B, ncu
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
(time timeout 900 python -m pytest tests/test_igemm_gpu.py tests/test_nn_gpu.py tests/test_keras_h5.py tests/test_full_size_gpu.py tests/test_adapters_gpu.py -m gpu -x -q -s) > $GOUT/c_pytest_nn.log 2>&1
echo "pytest rc=$?" >> $GOUT/c_pytest_nn.log
timeout 300 python tools/bench_nn.py > $GOUT/c_bench_nn_epi3.log 2>&1
CZ_EPI=2 timeout 300 python tools/bench_nn.py > $GOUT/c_bench_nn_epi2.log 2>&1
CZ_STAGES=4 timeout 300 python tools/bench_nn.py > $GOUT/c_bench_nn_epi3_4stages.log 2>&1
(time timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu) > $GOUT/c_bench_c3.log 2>&1
timeout 300 python bench.py --workload c2 --steps 12 --warmup 4 --no-cpu --no-secondary > $GOUT/c_bench_c2.log 2>&1
# ncu: launch list of a short c3 run + full capture of conv1 / conv2 of the new kernel at C=256 and C=128
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $GOUT/c_launches_c3.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary --sims 64 > $GOUT/c_ncu_c3_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_igemm3 -s 44 -c 2 -o $GOUT/c_igemm3_256 \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary --sims 64 > $GOUT/c_ncu_c3_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_igemm3 -s 44 -c 2 -o $GOUT/c_igemm3_128 \
    python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-secondary > $GOUT/c_ncu_c2_full.log 2>&1
ls -la $GOUT
