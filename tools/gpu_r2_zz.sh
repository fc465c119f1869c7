#!/bin/bash
# round 2, last GPU call: the cluster-4 experiment's test with an odd number of pair-tiles (the all-out-of-bounds tile)
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
(time timeout 150 python -m pytest tests/test_nn_gpu.py -m gpu -x -q -s -k cluster4) > $GOUT/zz_pytest.log 2>&1
echo "pytest rc=$?" >> $GOUT/zz_pytest.log
