#!/bin/bash
# round 2, GPU call U (8 GPUs): the driver's scaling command at N = 8 on the final tree
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 8 --steps 6 --warmup 3 > $GOUT/u_bench_c3_8gpu.log 2>&1
echo "rc=$?" >> $GOUT/u_bench_c3_8gpu.log
ls -la $GOUT
