#!/bin/bash
# round 2, GPU call 2: MMA issue microbenchmark 2, per-layer profile at b=8 and b=1, parity tests after the fixes
mkdir -p gpurun_out
O=gpurun_out
timeout 120 scripts/micro/mma_issue2.bin > $O/r2c2_mma_issue2.txt 2>&1; echo "microbench rc=$?"; cat $O/r2c2_mma_issue2.txt
VG_TOP=90 timeout 300 python scripts/profile_layers.py 128 8 > $O/r2c2_layers_b8.txt 2>&1; echo "layers b8 rc=$?"; head -75 $O/r2c2_layers_b8.txt
VG_TOP=40 timeout 300 python scripts/profile_layers.py 128 1 > $O/r2c2_layers_b1.txt 2>&1; echo "layers b1 rc=$?"; head -42 $O/r2c2_layers_b1.txt
timeout 900 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -s > $O/r2c2_pytest_parity.log 2>&1; echo "pytest parity rc=$?"; tail -8 $O/r2c2_pytest_parity.log
grep -E "gen block|gen layer|disc stage|step [0-9]+\^3|replay vs eager|worst single" $O/r2c2_pytest_parity.log | head -80
