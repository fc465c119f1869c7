#!/bin/bash
# round 2, GPU call Q (4 GPUs): the driver's scaling command at N = 4 on the final tree
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 4 --steps 6 --warmup 3 > $GOUT/q_bench_c3_4gpu.log 2>&1
echo "rc=$?" >> $GOUT/q_bench_c3_4gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 4 --workload c2 --steps 60 --warmup 5 --max-game-length 20 > $GOUT/q_bench_c2_4gpu_games_finish.log 2>&1
echo "rc=$?" >> $GOUT/q_bench_c2_4gpu_games_finish.log
ls -la $GOUT
