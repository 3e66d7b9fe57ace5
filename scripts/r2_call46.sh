#!/bin/bash
# per-layer profile at b=1 (what each rank of an 8-GPU job runs) + sliding-window phase profile
mkdir -p gpurun_out
VG_TOP=90 timeout 300 python scripts/profile_layers.py 128 1 > gpurun_out/r2c46_layers_b1.txt 2>&1; head -3 gpurun_out/r2c46_layers_b1.txt
VG_STITCH_PROFILE=1 timeout 300 python scripts/bench_configs.py sliding > gpurun_out/r2c46_sliding.txt 2>&1; tail -20 gpurun_out/r2c46_sliding.txt
