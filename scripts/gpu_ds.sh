#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for m in fwd dgrad; do timeout 100 python scripts/bench_conv.py $m 256-512; VG_TC_DSPLIT=0 timeout 100 python scripts/bench_conv.py $m 256-512; done
VG_TOP=200 timeout 400 python scripts/profile_layers.py 128 8 > gpurun_out/layers_b8.txt 2>&1; head -12 gpurun_out/layers_b8.txt
