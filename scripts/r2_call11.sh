#!/bin/bash
# round 2, GPU call 11: wgrad gather through LDG.128 + STS.128 (merged sector requests): parity + timing, tw-fold A/B
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv" > $O/r2c11_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2c11_pytest.log
VG_WG_NW=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv" > $O/r2c11_pytest_nw.log 2>&1; echo "pytest nw rc=$?"; tail -3 $O/r2c11_pytest_nw.log
echo "--- default"; timeout 300 python scripts/bench_conv.py wgrad > $O/r2c11_conv_wgrad.txt 2>&1; cat $O/r2c11_conv_wgrad.txt
echo "--- tw-fold"; VG_WG_NW=1 timeout 300 python scripts/bench_conv.py wgrad > $O/r2c11_conv_wgrad_nw.txt 2>&1; head -2 $O/r2c11_conv_wgrad_nw.txt
for b in "4,4" "4,2"; do echo "brick $b"; VG_WG_NW=1 VG_WG_BRICK=$b timeout 120 python scripts/bench_conv.py wgrad 16-16 2>&1 | tail -1;  VG_WG_NW=1 VG_WG_BRICK=$b timeout 120 python scripts/bench_conv.py wgrad 48-16 2>&1 | head -1; done
