#!/bin/bash
# round 2, GPU call 9: wgrad producer restructure (12 producer warps, incremental decode) parity + timing, tw-fold A/B
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv" > $O/r2c9_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2c9_pytest.log
echo "--- NW=0"; VG_WG_NW=0 timeout 300 python scripts/bench_conv.py wgrad > $O/r2c9_conv_wgrad_nw0.txt 2>&1; cat $O/r2c9_conv_wgrad_nw0.txt
echo "--- default (tw-fold)"; timeout 300 python scripts/bench_conv.py wgrad > $O/r2c9_conv_wgrad.txt 2>&1; head -2 $O/r2c9_conv_wgrad.txt
for b in "4,4" "8,2" "4,2"; do echo "brick $b"; VG_WG_BRICK=$b timeout 120 python scripts/bench_conv.py wgrad 16-16 2>&1 | tail -1;  VG_WG_BRICK=$b timeout 120 python scripts/bench_conv.py wgrad 48-16 2>&1 | head -1; done
