#!/bin/bash
# Freeze the working tree into snap_<name>/ so that a queued gpurun call runs exactly this state while development continues.
# usage: tools/snap.sh <name>   ->  gpurun -- 'cd snap_<name> && GOUT=../gpurun_out bash tools/<script>.sh'
set -e
name=$1
rm -rf "snap_$name"
mkdir -p "snap_$name"
tar -cf - --exclude='./.git' --exclude='./gpurun_out' --exclude='./snap_*' --exclude='.pytest_cache' \
    --exclude='./chinesechess-alphazero_b200/build' . | tar -xf - -C "snap_$name"
echo "snap_$name ready: $(du -sh snap_$name | cut -f1)"
