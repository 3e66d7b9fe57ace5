#!/bin/bash
# round 2, GPU call 43 (2 GPUs): final N=2 check: DP parity tests + bench (both arms)
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_dp_nccl.py -m gpu -q -x > $O/r2c43_pytest_dp.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r2c43_pytest_dp.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2c43_bench_n2.json 2> $O/r2c43_bench_n2.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('$O/r2c43_bench_n2.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d.get('sliding_window',{}).get('value'))"
