#!/bin/bash
# round 2, GPU call 25: vectorised weight pack + clip/Adam kernels: parity + bench at b=8 and b=1
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity_r2.py tests/test_gpu_train_step.py tests/test_gpu_monitor_ckpt.py -m gpu -q -x -k "not 128" > $O/r2c25_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2c25_pytest.log
for gb in 8 1; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sliding --global-batch $gb > $O/r2c25_bench_b$gb.json 2> $O/r2c25_bench_b$gb.err; echo "bench b=$gb rc=$?"; python -c "
import json;d=json.loads(open('$O/r2c25_bench_b$gb.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['roofline']['frac']);f=d['roofline']['families_ms_per_step'];print({k:f[k] for k in f if 'pack' in k or 'adam' in k})"
done
