#!/bin/bash
# round 2, GPU call 16: InstanceNorm backward with the reflect fold as an in-place pre-pass: parity + timing A/B
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity_r2.py -m gpu -q -x -k "instnorm or teacher or replay" > $O/r2c16_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2c16_pytest.log
echo "--- fold pre-pass (default)"; timeout 100 python scripts/bench_in.py 2>&1 | grep "^bwd" | tee $O/r2c16_in_fold.txt
echo "--- VG_IN_FOLD=0"; VG_IN_FOLD=0 timeout 100 python scripts/bench_in.py 2>&1 | grep "^bwd" | tee $O/r2c16_in_nofold.txt
