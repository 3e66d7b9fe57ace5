#!/bin/bash
# round 2, GPU call 3: unrolled d-march issue, fused weight repack, order-exact stitcher, monitor / checkpoint tests
mkdir -p gpurun_out
O=gpurun_out
timeout 120 scripts/micro/mma_issue3.bin > $O/r2c3_mma_issue3.txt 2>&1; echo "microbench3 rc=$?"; cat $O/r2c3_mma_issue3.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv3d or clip_adam" > $O/r2c3_pytest_conv.log 2>&1; echo "conv pytest rc=$?"; tail -3 $O/r2c3_pytest_conv.log
timeout 900 python -m pytest tests/test_gpu_monitor_ckpt.py tests/test_gpu_golden_stitch.py -m gpu -q > $O/r2c3_pytest_monitor.log 2>&1; echo "monitor pytest rc=$?"; tail -30 $O/r2c3_pytest_monitor.log
timeout 300 python scripts/bench_conv.py fwd > $O/r2c3_conv_fwd.txt 2>&1; head -4 $O/r2c3_conv_fwd.txt
timeout 300 python scripts/bench_conv.py dgrad > $O/r2c3_conv_dgrad.txt 2>&1; head -4 $O/r2c3_conv_dgrad.txt
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_kernels.py > $O/r2c3_pytest_rest.log 2>&1; echo "rest pytest rc=$?"; tail -5 $O/r2c3_pytest_rest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2c3_bench.json 2> $O/r2c3_bench.err; echo "bench rc=$?"; cut -c1-300 $O/r2c3_bench.json; tail -3 $O/r2c3_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c3_bench.json'))
print(d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d.get('sliding_window'))
print(d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['families_ms_per_step'])
PY
