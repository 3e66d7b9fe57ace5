#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"in_bwd_apply|in_bwd_partial|in_apply_kernel|in_stats_partial" -s 8 -c 4 -o gpurun_out/in_kernels -f python scripts/bench_in.py one > gpurun_out/ncu_in.log 2>&1
tail -3 gpurun_out/ncu_in.log
