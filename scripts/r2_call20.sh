#!/bin/bash
# round 2, GPU call 20: 4-chain forward + weight-gradient sibling streams: parity + timing at b=8 and b=1, A/B
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_train_step.py tests/test_gpu_monitor_ckpt.py tests/test_gpu_vnet_si.py -m gpu -q -x -k "not 128" > $O/r2c20_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2c20_pytest.log
for wg in 1 0; do
  for gb in 8 1; do
    VG_WG_STREAM=$wg timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sliding --global-batch $gb > $O/r2c20_bench_wg${wg}_b$gb.json 2> $O/r2c20_bench_wg${wg}_b$gb.err; echo "bench wgstream=$wg b=$gb rc=$?"; python -c "
import json;d=json.loads(open('$O/r2c20_bench_wg${wg}_b$gb.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['peak_mem_gib'])"
  done
done
