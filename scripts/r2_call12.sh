#!/bin/bash
# round 2, GPU call 12 (2 GPUs): vg_comm + in-graph overlapped all-reduce: N=2 parity tests, bench A/B
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_dp_nccl.py -m gpu -q -x -s > $O/r2c12_pytest_dp.log 2>&1; echo "pytest rc=$?"; grep -n "N=2\|passed\|failed\|Error" $O/r2c12_pytest_dp.log | head -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-sliding > $O/r2c12_bench_n2.json 2> $O/r2c12_bench_n2.err; echo "bench rc=$?"; tail -3 $O/r2c12_bench_n2.err; python -c "
import json;d=json.loads(open('$O/r2c12_bench_n2.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['launch_mode'],d.get('comm'))"
VG_GRAPH_COMM=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-sliding > $O/r2c12_bench_n2_blocking.json 2> $O/r2c12_bench_n2_blocking.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('$O/r2c12_bench_n2_blocking.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['launch_mode'])"
