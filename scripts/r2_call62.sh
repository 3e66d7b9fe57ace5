#!/bin/bash
# round 2, GPU call 62: capture layouts at b = 1 on one GPU (what the split costs without any exchange); new tests
O=gpurun_out/r2c62.txt
: > $O
timeout 300 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_kernels.py -m gpu -q -k "graph_replay or gradient_scale" 2>&1 | tail -1 >> $O
for sp in 0 1 2; do
VG_GRAPH_SPLIT=$sp timeout 300 python bench.py --steps 10 --warmup 3 --global-batch 1 --no-cpu-baseline --no-sliding > gpurun_out/r2c62_b1_split$sp.json 2> /dev/null
python -c "
import json;d=json.loads(open('gpurun_out/r2c62_b1_split$sp.json').read().strip().splitlines()[-1]);print('b1 VG_GRAPH_SPLIT=$sp', d['ms_per_step'],d['value'])" >> $O
done
cat $O
