#!/bin/bash
# round 2, GPU call Y: 4-CTA clusters, weights multicast to both CTA pairs (CZ_CLUSTER4=1, experiment)
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
(time CZ_CLUSTER4=1 CZ_EPI=3 timeout 240 python -m pytest tests/test_nn_gpu.py -m gpu -x -q -k "small_batch or restatement") > $GOUT/y_pytest_c4.log 2>&1
echo "pytest rc=$?" >> $GOUT/y_pytest_c4.log
AB_SHAPES=c3 AB_ONLY="skip default,conv2 on igemm3,cluster4" timeout 500 python tools/ab_nn.py 2 > $GOUT/y_ab_nn.log 2>&1
ls -la $GOUT
