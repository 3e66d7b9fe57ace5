#!/bin/bash
# round 2, GPU call 4: full GPU suite with the default path, DMMAX=192 A/B, sliding-window phase profile
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r2c4_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -6 $O/r2c4_pytest_all.log
VG_TC_DMMAX=192 timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_parity_r2.py tests/test_gpu_kernels.py -m gpu -q -k "not full_size" > $O/r2c4_pytest_dm192.log 2>&1; echo "pytest dm192 rc=$?"; tail -6 $O/r2c4_pytest_dm192.log
VG_TC_DMMAX=192 timeout 300 python scripts/bench_conv.py fwd > $O/r2c4_conv_fwd_dm192.txt 2>&1
VG_TC_DMMAX=192 timeout 300 python scripts/bench_conv.py dgrad > $O/r2c4_conv_dgrad_dm192.txt 2>&1
paste -d'|' gpurun_out/r2c3_conv_fwd.txt $O/r2c4_conv_fwd_dm192.txt | cut -c1-62,118-180 | head -14
paste -d'|' gpurun_out/r2c3_conv_dgrad.txt $O/r2c4_conv_dgrad_dm192.txt | cut -c1-62,118-180 | head -14
VG_STITCH_PROFILE=1 timeout 600 python scripts/bench_configs.py sliding > $O/r2c4_sliding.txt 2>&1; echo "sliding rc=$?"; tail -8 $O/r2c4_sliding.txt | cut -c1-400
VG_TC_DMMAX=192 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sliding > $O/r2c4_bench_dm192.json 2> $O/r2c4_bench_dm192.err; echo "bench rc=$?"; cut -c1-200 $O/r2c4_bench_dm192.json
