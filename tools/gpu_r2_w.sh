#!/bin/bash
# round 2, GPU call W: which part of conv2's fp32 epilogue slows its mainloop?  (perf-only switches, results are wrong by design)
set -x
GOUT=${GOUT:-gpurun_out}; mkdir -p $GOUT
AB_SHAPES=c3 AB_ONLY="skip default,conv2 on igemm3,dbg" timeout 900 python tools/ab_nn.py 2 > $GOUT/w_ab_nn.log 2>&1
ls -la $GOUT
