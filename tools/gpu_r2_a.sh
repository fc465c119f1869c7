#!/bin/bash
# round 2, GPU call A: full GPU test tier, smoke, baseline bench (c3 + secondary + CPU arm), launch list + ncu of the C=128 conv
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $GOUT/a_smi.txt
(time timeout 1500 python -m pytest tests -m gpu -x -q -s) > $GOUT/a_pytest.log 2>&1
echo "pytest rc=$?" >> $GOUT/a_pytest.log
timeout 300 python __graft_entry__.py smoke > $GOUT/a_smoke.log 2>&1
(time timeout 900 python bench.py --steps 4 --warmup 3) > $GOUT/a_bench_c3.log 2>&1
(time timeout 600 python bench.py --impl reference --steps 6 --warmup 1) > $GOUT/a_bench_ref.log 2>&1
timeout 300 python tools/bench_nn.py > $GOUT/a_bench_nn.log 2>&1
timeout 300 python tools/bench_tree.py > $GOUT/a_bench_tree.log 2>&1
# launch list of a short c2 run, then ONE full capture of the C=128 conv kernel inside it
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $GOUT/a_launches_c2.csv \
    python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-secondary > $GOUT/a_ncu_c2_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_igemm2 -s 40 -c 2 -o $GOUT/a_igemm2_128 \
    python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-secondary > $GOUT/a_ncu_c2_full.log 2>&1
ls -la $GOUT
# does the driver run WHILE conditional graph nodes? (expect "0 no error")
./tools/cond_graph_probe > $GOUT/a_cond_graph.log 2>&1
# does overlapping the tree kernels with the network pay at c2? (two-range pipeline forced on)
CZ_FORCE_PIPELINE=1 timeout 300 python bench.py --workload c2 --steps 10 --warmup 4 --no-cpu --no-secondary > $GOUT/a_bench_c2_pipelined.log 2>&1
timeout 300 python bench.py --workload c2 --steps 10 --warmup 4 --no-cpu --no-secondary > $GOUT/a_bench_c2_seq.log 2>&1
