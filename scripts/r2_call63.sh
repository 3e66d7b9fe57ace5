#!/bin/bash
# round 2, GPU call 63 (4 GPUs): capture layout A/B at N = 4 (b = 2 per GPU): per-sweep (three graphs) vs two graphs
O=gpurun_out
for lay in per-sweep two; do
VG_GRAPH_LAYOUT=$lay timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline --no-sliding > $O/r2c63_n4_$lay.json 2> $O/r2c63_n4_$lay.err; echo "rc=$?"
python -c "
import json;d=json.loads(open('$O/r2c63_n4_$lay.json').read().strip().splitlines()[-1]);print('$lay', d['ms_per_step'],d['value'],d['e2e']['value'])"
done
