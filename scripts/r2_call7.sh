#!/bin/bash
# round 2, GPU call 7: wgrad tw-fold (N = 144) parity + timing A/B, vnet gen_SI tests
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_vnet_si.py -m gpu -q -x > $O/r2c7_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r2c7_pytest.log
for cfg in "16-16" "48-16"; do
  VG_WG_NW=0 timeout 120 python scripts/bench_conv.py wgrad $cfg 2>&1 | tail -1
  VG_DEBUG=1 timeout 120 python scripts/bench_conv.py wgrad $cfg 2>&1 | grep -v "^\[wgrad_tc\]" | tail -1
  VG_DEBUG=1 timeout 120 python scripts/bench_conv.py wgrad $cfg 2>&1 | grep "^\[wgrad_tc\]" | tail -1
  for b in "8,4" "4,4" "8,2" "4,2" "2,2"; do echo "brick $b"; VG_WG_BRICK=$b timeout 120 python scripts/bench_conv.py wgrad $cfg 2>&1 | tail -1; done
done
timeout 300 python scripts/bench_conv.py wgrad > $O/r2c7_conv_wgrad.txt 2>&1; cat $O/r2c7_conv_wgrad.txt
