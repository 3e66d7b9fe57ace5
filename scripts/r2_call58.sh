#!/bin/bash
# round 2, GPU call 58 (8 GPUs): N=8 bench (does the 3-graph + eager all-reduce step run and scale at b=1 per GPU)
O=gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 > $O/r2c58_bench_n8.json 2> $O/r2c58_bench_n8.err; echo "bench rc=$?"; tail -3 $O/r2c58_bench_n8.err; python -c "
import json;d=json.loads(open('$O/r2c58_bench_n8.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d.get('sliding_window',{}).get('value'))"
