#!/bin/bash
# round 2, GPU call V: the producer warp's idle lanes prefetch the next tile's input rows into L2 (CZ_ACT_PREFETCH=1)
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
(time CZ_EPI=3 CZ_ACT_PREFETCH=1 timeout 600 python -m pytest tests/test_nn_gpu.py -m gpu -x -q) > $GOUT/v_pytest_pf.log 2>&1
echo "pytest rc=$?" >> $GOUT/v_pytest_pf.log
AB_SHAPES=c3 AB_ONLY="skip default,pf" timeout 900 python tools/ab_nn.py 3 > $GOUT/v_ab_nn.log 2>&1
ls -la $GOUT
