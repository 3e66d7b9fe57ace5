#!/bin/bash
# round 2, GPU call 22 (2 GPUs): 3-graph capture + side streams + vg_comm at bench scale; N=2 parity tests
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_dp_nccl.py -m gpu -q -x -s > $O/r2c22_pytest_dp.log 2>&1; echo "pytest rc=$?"; grep -n "N=2\|passed\|failed\|Error" $O/r2c22_pytest_dp.log | head -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2c22_bench_n2.json 2> $O/r2c22_bench_n2.err; echo "bench rc=$?"; tail -2 $O/r2c22_bench_n2.err; python -c "
import json;d=json.loads(open('$O/r2c22_bench_n2.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['launch_mode'],d.get('comm'),d.get('sliding_window'))"
