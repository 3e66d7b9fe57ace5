#!/bin/bash
# round 2, GPU call 59: first InstanceNorm pass walks the samples in descending order (L2 reuse across the pass pair)
O=gpurun_out/r2c59.txt
: > $O
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_vnet_si.py -m gpu -q -x -k "instnorm or batchnorm or norm" 2>&1 | tail -1 >> $O
timeout 200 python scripts/bench_in.py >> $O 2>&1
for gb in 8 2; do
timeout 300 python bench.py --steps 5 --warmup 3 --global-batch $gb --no-cpu-baseline --no-sliding > gpurun_out/r2c59_bench_b$gb.json 2> gpurun_out/r2c59_bench_b$gb.err
python -c "
import json;d=json.loads(open('gpurun_out/r2c59_bench_b$gb.json').read().strip().splitlines()[-1]);print('b$gb', d['ms_per_step'],d['value'],d['roofline']['frac'], {k:(round(v['frac'],3),v['ms_per_step']) for k,v in d['roofline_other_kernels'].items() if k.startswith('instnorm')})" >> $O
done
cat $O
