#!/bin/bash
# round 2, GPU call 55: deterministic skeleton backward (fixed-point ring), pad_noise / upsample_concat / cin1 forward changes
O=gpurun_out/r2c55.txt
: > $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/r2c55_pytest_kernels.log 2>&1; echo "kernels rc=$?" >> $O; tail -2 gpurun_out/r2c55_pytest_kernels.log >> $O
timeout 120 python scripts/diag_skel.py 100 2x32x32x32 tanh >> $O 2>&1
timeout 120 python scripts/diag_skel.py 30 1x128x128x128 tanh >> $O 2>&1
for i in 1 2 3 4; do timeout 200 python -m pytest "tests/test_gpu_parity_r2.py::test_graph_replay_equals_eager" -m gpu -q 2>&1 | tail -1 >> $O; done
VG_STREAMS=0 timeout 250 python scripts/diag_nondet.py 12 2>&1 | grep -E "deviating" >> $O
python scripts/bench_skel.py 8 >> $O 2>&1
python scripts/bench_skel.py 1 >> $O 2>&1
for m in fwd; do for s in 1-16 1-64; do python scripts/bench_conv.py $m $s >> $O 2>&1; VG_CIN1_PF=0 python scripts/bench_conv.py $m $s >> $O 2>&1; done; done
timeout 600 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_golden_stitch.py -m gpu -q -x 2>&1 | tail -2 >> $O
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sliding > gpurun_out/r2c55_bench_b8.json 2> gpurun_out/r2c55_bench_b8.err; python -c "
import json;d=json.loads(open('gpurun_out/r2c55_bench_b8.json').read().strip().splitlines()[-1]);print('b8', d['ms_per_step'],d['value'],d['roofline']['frac'])" >> $O
cat $O
