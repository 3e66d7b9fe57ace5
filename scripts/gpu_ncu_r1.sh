#!/bin/bash
# ncu evidence for a round (run under gpurun, ONE GPU):
#   1. `--set full` capture of the dominant kernel (tc_conv, fwd 48->16 at 8x130^3) and of the 16->16 wgrad  -> .ncu-rep
#   2. launch list of ONE train step of bench.py's workload (scripts/one_step.py): every launch costs ~55 ms under ncu,
#      so the full bench command (7 steps, 14 400 launches) would be ~13 GPU-minutes for the same per-step list
# digest here with: python scripts/ncu_digest.py gpurun_out/<name>.ncu-rep ; python scripts/launch_summary.py gpurun_out/launches_step.csv 1
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv -s 2 -c 1 -o gpurun_out/fwd_48-16 -f python scripts/bench_conv.py fwd 48-16 > gpurun_out/ncu_fwd.log 2>&1; tail -1 gpurun_out/ncu_fwd.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc -s 2 -c 1 -o gpurun_out/wg_16-16 -f python scripts/bench_conv.py wgrad 16-16 > gpurun_out/ncu_wg.log 2>&1; tail -1 gpurun_out/ncu_wg.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches_step.csv python scripts/one_step.py > gpurun_out/one_step.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/launches_step.csv
