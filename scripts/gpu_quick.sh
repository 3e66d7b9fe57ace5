#!/bin/bash
# quick loop: kernel parity tests + per-layer in-step profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 400 python scripts/profile_layers.py 128 8 > gpurun_out/layers_b8.txt 2>&1; head -45 gpurun_out/layers_b8.txt
