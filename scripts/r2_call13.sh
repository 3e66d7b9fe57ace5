#!/bin/bash
# round 2, GPU call 13 (1 GPU): split-graph capture == eager, vnet_si, full regression of the step tests, bench
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_vnet_si.py tests/test_gpu_train_step.py tests/test_gpu_monitor_ckpt.py -m gpu -q -x -k "not 128" > $O/r2c13_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2c13_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sliding > $O/r2c13_bench.json 2> $O/r2c13_bench.err; echo "bench rc=$?"; cut -c1-400 $O/r2c13_bench.json
