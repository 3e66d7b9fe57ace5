#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/tcdbg.txt gpurun_out/convbench.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "conv3d" > gpurun_out/pytest_k.log 2>&1; tail -3 gpurun_out/pytest_k.log
for dbg in 0 1 2 3; do for sh in 16-16 48-16; do echo "VG_TC_DEBUG=$dbg" >> gpurun_out/tcdbg.txt; VG_TC_DEBUG=$dbg timeout 100 python scripts/bench_conv.py fwd $sh >> gpurun_out/tcdbg.txt 2>&1; done; done
cat gpurun_out/tcdbg.txt
for m in fwd dgrad; do timeout 300 python scripts/bench_conv.py $m >> gpurun_out/convbench.txt 2>&1; done
cat gpurun_out/convbench.txt
