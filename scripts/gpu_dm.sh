#!/bin/bash
# d-march validation: conv parity tests, isolated conv timings (old loop / BD 4 / BD 8), per-layer in-step profile
mkdir -p gpurun_out; rm -f gpurun_out/convbench_dm.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k conv3d > gpurun_out/pytest_conv.log 2>&1; tail -8 gpurun_out/pytest_conv.log
for cfg in "VG_TC_BD=4" "VG_TC_BD=8"; do
  for m in fwd dgrad; do echo "== $cfg $m" >> gpurun_out/convbench_dm.txt; env $cfg timeout 300 python scripts/bench_conv.py $m >> gpurun_out/convbench_dm.txt 2>&1; done
done
cat gpurun_out/convbench_dm.txt
VG_TOP=200 timeout 400 python scripts/profile_layers.py 128 8 > gpurun_out/layers_b8.txt 2>&1; head -40 gpurun_out/layers_b8.txt
