#!/bin/bash
# round 2, GPU call 57: ncu evidence of the final state: --set full of the skeleton backward routing kernel, launch list of ONE step (b = 8)
O=gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:skel_bwd_route_march -s 5 -c 1 -o $O/r2_skel_route_march -f python scripts/bench_skel.py 8 > $O/ncu_skel.log 2>&1; tail -1 $O/ncu_skel.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $O/r2_launches_step_final.csv python scripts/one_step.py > $O/one_step_final.log 2>&1; echo "launch list rc=$?"; wc -l $O/r2_launches_step_final.csv
ls -la $O/*.ncu-rep | tail -3
