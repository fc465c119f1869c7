#!/bin/bash
# round 2, GPU call F (2 GPUs): the N > 1 paths — bench.py under torchrun (in-loop NCCL gather of the record rings, rank-0 decode)
# and the data-parallel self_play.start entry
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
nvidia-smi -L > $GOUT/f_gpus.txt
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $GOUT/f_pytest.log 2>&1
echo "pytest rc=$?" >> $GOUT/f_pytest.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3) > $GOUT/f_bench_c3_2gpu.log 2>&1
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c2 --steps 30 --warmup 5 --max-game-length 10 --no-secondary) > $GOUT/f_bench_c2_2gpu_games_finish.log 2>&1
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/run_selfplay_dp.py /tmp/cz_dp 600) > $GOUT/f_selfplay_dp.log 2>&1
ls /tmp/cz_dp/play_data | head -3 >> $GOUT/f_selfplay_dp.log
(time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 4 --warmup 1 --no-secondary) > $GOUT/f_bench_ref_2gpu.log 2>&1
ls -la $GOUT
