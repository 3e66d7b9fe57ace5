#!/bin/bash
# round 2, GPU call 47 (4 GPUs): N=4 bench (does the 3-graph + eager all-reduce step run and scale at b=2 per GPU)
O=gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 5 --warmup 3 > $O/r2c47_bench_n4.json 2> $O/r2c47_bench_n4.err; echo "bench rc=$?"; tail -3 $O/r2c47_bench_n4.err; python -c "
import json;d=json.loads(open('$O/r2c47_bench_n4.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d.get('sliding_window',{}).get('value'))"
