#!/bin/bash
# round 2, GPU call 38: every loss swept as its independent parts on separate streams: parity + bench A/B at b=8 and b=1
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_train_step.py tests/test_gpu_monitor_ckpt.py tests/test_gpu_vnet_si.py tests/test_gpu_resnet.py -m gpu -q -x -k "not 128" > $O/r2c38_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2c38_pytest.log
for sp in 1 0; do
  for gb in 8 1; do
    VG_SPLIT_SWEEPS=$sp timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sliding --global-batch $gb > $O/r2c38_bench_sp${sp}_b$gb.json 2>/dev/null; python -c "
import json;d=json.loads(open('$O/r2c38_bench_sp${sp}_b$gb.json').read().strip().splitlines()[-1]);print('split=$sp b=$gb', d['ms_per_step'],d['value'],d['peak_mem_gib'])"
  done
done
