#!/bin/bash
# ncu evidence of round 2 (run under gpurun, ONE GPU): --set full captures of the dominant conv kernel, the 16->16 weight gradient and
# the three InstanceNorm backward kernels (fold pre-pass, folded partial, folded apply) + the launch list of ONE train step.
# digest: python scripts/ncu_digest.py gpurun_out/<name>.ncu-rep ; python scripts/launch_summary.py gpurun_out/r2_launches_step.csv 1
mkdir -p gpurun_out
O=gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv -s 2 -c 1 -o $O/r2_fwd_48-16 -f python scripts/bench_conv.py fwd 48-16 > $O/ncu_fwd.log 2>&1; tail -1 $O/ncu_fwd.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc -s 2 -c 1 -o $O/r2_wg_16-16 -f python scripts/bench_conv.py wgrad 16-16 > $O/ncu_wg.log 2>&1; tail -1 $O/ncu_wg.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"in_bwd_partial_folded|in_bwd_apply_sp|in_fold_inplace" -s 6 -c 3 -o $O/r2_in_bwd_c16 -f python scripts/bench_in.py one > $O/ncu_in.log 2>&1; tail -1 $O/ncu_in.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $O/r2_launches_step.csv python scripts/one_step.py > $O/one_step.log 2>&1; echo "launch list rc=$?"; wc -l $O/r2_launches_step.csv
