#!/bin/bash
# round 2, GPU call 56: thinner tc_conv bricks on small grids, block-reduced cout1 weight gradient
O=gpurun_out/r2c56.txt
: > $O
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv3d" > gpurun_out/r2c56_pytest_conv.log 2>&1; echo "conv tests rc=$?" >> $O; tail -2 gpurun_out/r2c56_pytest_conv.log >> $O
for gb in 1 8; do
timeout 300 python bench.py --steps 5 --warmup 3 --global-batch $gb --no-cpu-baseline --no-sliding > gpurun_out/r2c56_bench_b$gb.json 2> gpurun_out/r2c56_bench_b$gb.err
python -c "
import json;d=json.loads(open('gpurun_out/r2c56_bench_b$gb.json').read().strip().splitlines()[-1]);print('b$gb', d['ms_per_step'],d['value'],d['roofline']['frac'])" >> $O
done
VG_TC_SMALLGRID=0 timeout 300 python bench.py --steps 5 --warmup 3 --global-batch 1 --no-cpu-baseline --no-sliding > gpurun_out/r2c56_bench_b1_off.json 2> /dev/null
python -c "
import json;d=json.loads(open('gpurun_out/r2c56_bench_b1_off.json').read().strip().splitlines()[-1]);print('b1 smallgrid off', d['ms_per_step'],d['value'],d['roofline']['frac'])" >> $O
VG_TOP=60 timeout 300 python scripts/profile_layers.py 128 1 > gpurun_out/r2c56_layers_b1.txt 2>&1
cat $O
