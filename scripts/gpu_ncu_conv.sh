#!/bin/bash
# usage: gpu_ncu_conv.sh <mode> <shape> [kernel regex]
mkdir -p gpurun_out
mode=$1; sh=$2; rx=${3:-tc_conv}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -o gpurun_out/${mode}_$sh -f python scripts/bench_conv.py $mode $sh > gpurun_out/ncu_${mode}_$sh.log 2>&1
tail -2 gpurun_out/ncu_${mode}_$sh.log
