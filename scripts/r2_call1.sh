#!/bin/bash
# round 2, GPU call 1: validate the lean d-march issue loop (VG_TC_DMLEAN=1), run the new parity tests, baseline timings
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L; nproc; free -g | head -2
VG_TC_DMLEAN=1 timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "conv3d" > $O/r2c1_pytest_dmlean.log 2>&1; echo "dmlean conv pytest rc=$?"; tail -3 $O/r2c1_pytest_dmlean.log
for m in fwd dgrad; do
  timeout 300 python scripts/bench_conv.py $m > $O/r2c1_conv_${m}_default.txt 2>&1
  VG_TC_DMLEAN=1 timeout 300 python scripts/bench_conv.py $m > $O/r2c1_conv_${m}_lean.txt 2>&1
done
timeout 300 python scripts/bench_conv.py wgrad > $O/r2c1_conv_wgrad.txt 2>&1
paste -d'|' $O/r2c1_conv_fwd_default.txt $O/r2c1_conv_fwd_lean.txt | cut -c1-200
timeout 1500 python -m pytest tests -m gpu -q -s > $O/r2c1_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -15 $O/r2c1_pytest_all.log
grep -E "gen block|gen layer|disc stage|step [0-9]+\^3|128\^3|replay vs eager|worst single" $O/r2c1_pytest_all.log | head -120
timeout 400 python bench.py --steps 3 --warmup 3 --no-sliding --no-cpu-baseline > $O/r2c1_bench_default.json 2> $O/r2c1_bench_default.err; echo "bench rc=$?"; cut -c1-400 $O/r2c1_bench_default.json
VG_TC_DMLEAN=1 timeout 400 python bench.py --steps 3 --warmup 3 --no-sliding --no-cpu-baseline > $O/r2c1_bench_lean.json 2> $O/r2c1_bench_lean.err; echo "bench lean rc=$?"; cut -c1-400 $O/r2c1_bench_lean.json
