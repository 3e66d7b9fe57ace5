#!/bin/bash
# round 2, GPU call 35: vnet teacher-forced gradients (both variants) + default bench + b=1 bench
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_vnet_si.py tests/test_gpu_vnet.py -m gpu -q -x -s > $O/r2c35_pytest_vnet.log 2>&1; echo "pytest rc=$?"; grep "vnet gen_\|passed\|failed" $O/r2c35_pytest_vnet.log | cut -c1-200 | tail -24
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sliding --global-batch 1 > $O/r2c35_bench_b1.json 2>/dev/null; python -c "
import json;d=json.loads(open('$O/r2c35_bench_b1.json').read().strip().splitlines()[-1]);print('b=1', d['ms_per_step'],d['value'])"
