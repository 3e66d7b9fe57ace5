#!/bin/bash
# One gpurun call: GPU parity tests, bench, ncu launch list, isolated conv bench, ncu --set full of the top conv kernels.
mkdir -p gpurun_out
rm -f gpurun_out/convbench.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python scripts/profile_step.py 128 1 2 > gpurun_out/profile_step.log 2>&1; echo "ncu rc=$?"
python scripts/launch_summary.py gpurun_out/launches.csv 2 > gpurun_out/launch_summary.txt; head -30 gpurun_out/launch_summary.txt
bash scripts/gpu_convbench.sh
bash scripts/gpu_ncu_conv.sh fwd 48-16 tc_conv
bash scripts/gpu_ncu_wgrad.sh
