#!/bin/bash
mkdir -p gpurun_out
for m in fwd dgrad wgrad; do timeout 300 python scripts/bench_conv.py $m >> gpurun_out/convbench.txt 2>&1; done
cat gpurun_out/convbench.txt
