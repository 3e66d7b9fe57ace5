#!/bin/bash
# ncu evidence for the round: full capture of the dominant kernel (tc_conv, d-march, fwd 48->16 at 8x130^3), of the 16->16 wgrad,
# and the launch list of one bench step (eager launches so that every kernel is a separate ncu record)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv -s 2 -c 1 -o gpurun_out/fwd_48-16_dm -f python scripts/bench_conv.py fwd 48-16 > gpurun_out/ncu_fwd.log 2>&1; tail -1 gpurun_out/ncu_fwd.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc -s 2 -c 1 -o gpurun_out/wg_16-16 -f python scripts/bench_conv.py wgrad 16-16 > gpurun_out/ncu_wg.log 2>&1; tail -1 gpurun_out/ncu_wg.log
VG_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 19500 -c 6600 --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/launches_step.csv
