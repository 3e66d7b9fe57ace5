#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/convbench.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "conv3d or instnorm" > gpurun_out/pytest_k.log 2>&1; tail -3 gpurun_out/pytest_k.log
timeout 300 python scripts/bench_in.py > gpurun_out/bench_in.txt 2>&1; cat gpurun_out/bench_in.txt
for m in fwd dgrad; do timeout 300 python scripts/bench_conv.py $m >> gpurun_out/convbench.txt 2>&1; done
timeout 100 python scripts/bench_conv.py wgrad 1-16 >> gpurun_out/convbench.txt 2>&1
cat gpurun_out/convbench.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
