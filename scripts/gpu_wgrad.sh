#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k conv3d -x > gpurun_out/wg_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/wg_pytest.log
tail -15 gpurun_out/wg_pytest.log
VG_DEBUG=1 timeout 200 python scripts/bench_conv.py wgrad > gpurun_out/wg_bench.txt 2>&1; echo "bench rc=$?" >> gpurun_out/wg_bench.txt
cat gpurun_out/wg_bench.txt
