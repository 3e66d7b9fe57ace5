#!/bin/bash
# round 2, GPU call 19: side-stream concurrency (forward chains, backward sweeps): parity + timing at b=8 and b=1, A/B
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_train_step.py tests/test_gpu_monitor_ckpt.py tests/test_gpu_vnet.py tests/test_gpu_vnet_si.py -m gpu -q -x -k "not 128" > $O/r2c19_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2c19_pytest.log
for st in 1 0; do
  for gb in 8 1; do
    VG_STREAMS=$st timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sliding --global-batch $gb > $O/r2c19_bench_st${st}_b$gb.json 2> $O/r2c19_bench_st${st}_b$gb.err; echo "bench streams=$st b=$gb rc=$?"; python -c "
import json;d=json.loads(open('$O/r2c19_bench_st${st}_b$gb.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['peak_mem_gib'])"
  done
done
