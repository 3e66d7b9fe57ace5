#!/bin/bash
mkdir -p gpurun_out
for sh in 48-16 256-512; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc -s 2 -c 1 -o gpurun_out/wg_$sh -f python scripts/bench_conv.py wgrad $sh > gpurun_out/ncu_wg_$sh.log 2>&1
tail -2 gpurun_out/ncu_wg_$sh.log
done
