#!/bin/bash
# round 2, GPU call G (2 GPUs): N = 2 bench after the device-index fix, fused apply+wave / block scan parity, split-producer A/B
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
(time timeout 900 python -m pytest tests/test_search.py tests/test_selfplay.py tests/test_arena.py tests/test_games_golden.py tests/test_adapters_gpu.py tests/test_full_size_gpu.py tests/test_uci.py -m gpu -x -q) > $GOUT/g_pytest.log 2>&1
echo "pytest rc=$?" >> $GOUT/g_pytest.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3) > $GOUT/g_bench_c3_2gpu.log 2>&1
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c2 --steps 30 --warmup 5 --max-game-length 10 --no-secondary) > $GOUT/g_bench_c2_2gpu_games_finish.log 2>&1
export CUDA_VISIBLE_DEVICES=0
for v in 1 2; do
  timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu --no-secondary > $GOUT/g_bench_c2_base_$v.log 2>&1
  CZ_SPLIT_PROD=1 timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu --no-secondary > $GOUT/g_bench_c2_split_$v.log 2>&1
done
CZ_SPLIT_PROD=1 timeout 300 python -m pytest tests/test_igemm_gpu.py tests/test_nn_gpu.py -m gpu -x -q > $GOUT/g_pytest_split.log 2>&1
timeout 300 python tools/bench_tree.py > $GOUT/g_bench_tree.log 2>&1
ls -la $GOUT
