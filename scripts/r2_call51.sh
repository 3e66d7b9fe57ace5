#!/bin/bash
# bisect test_graph_replay_equals_eager[False] by environment switches
O=gpurun_out/r2c51_bisect.txt
: > $O
run() { echo "=== $*" >> $O; env "$@" timeout 200 python -m pytest "tests/test_gpu_parity_r2.py::test_graph_replay_equals_eager" -m gpu -q -s -k False 2>&1 | grep -E "replay vs eager|passed|failed|Error" >> $O; }
run A=1
run VG_SKEL_BWD=tile
run VG_SMALL=0
run VG_BIAS_SINKS=0
run VG_STATS_REUSE=0
run VG_STREAMS=0
run VG_WG_STREAM=0
run VG_SKEL=tile
cat $O
