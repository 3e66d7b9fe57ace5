#!/bin/bash
# round 2, GPU call 48: k1 s1 dgrad on the streaming kernel, skeleton tie-rule test, step time at b=8 and b=1
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv3d or skel" > $O/r2c48_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2c48_pytest.log
for gb in 8 1; do
timeout 300 python bench.py --steps 5 --warmup 3 --global-batch $gb --no-cpu-baseline --no-sliding > $O/r2c48_bench_b$gb.json 2> $O/r2c48_bench_b$gb.err; echo "bench b=$gb rc=$?"
python -c "
import json;d=json.loads(open('$O/r2c48_bench_b$gb.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac'],d.get('roofline_other_kernels'))"
done
