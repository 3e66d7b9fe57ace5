#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden_stitch.py -m gpu -q -k "skel or losses or stitch or upsample" > gpurun_out/pytest_skel.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_skel.log
timeout 300 python scripts/bench_configs.py losses > gpurun_out/configs_losses.jsonl 2> gpurun_out/configs.err; grep -E '"S": (128|256)' gpurun_out/configs_losses.jsonl | grep -E '"iters": (15|50)|cycle_seg'; tail -2 gpurun_out/configs.err
VG_SKEL=tile timeout 300 python scripts/bench_configs.py losses 2>/dev/null | grep -E '"S": 128' | grep -E '"iters": 15|cycle_seg'
timeout 300 python scripts/bench_configs.py sliding 2>&1 | tail -2
