#!/bin/bash
# round 2, GPU call 64 (2 GPUs): per-sweep layout with two update graphs: replay == eager on one GPU, DP parity tests and bench on two
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -k "graph_replay" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_dp_nccl.py -m gpu -q -x > $O/r2c64_pytest_dp.log 2>&1; echo "dp pytest rc=$?"; tail -1 $O/r2c64_pytest_dp.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-sliding > $O/r2c64_bench_n2.json 2> $O/r2c64_bench_n2.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('$O/r2c64_bench_n2.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'])"
