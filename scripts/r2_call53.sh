#!/bin/bash
O=gpurun_out/r2c53_skel_diag.txt
: > $O
timeout 120 python scripts/diag_skel.py 200 2x32x32x32 tanh >> $O 2>&1
VG_SKEL_BWD=tile timeout 120 python scripts/diag_skel.py 200 2x32x32x32 tanh >> $O 2>&1
timeout 120 python scripts/diag_skel.py 100 2x32x32x32 quant >> $O 2>&1
timeout 120 python scripts/diag_skel.py 100 2x32x32x32 rand >> $O 2>&1
timeout 120 python scripts/diag_skel.py 50 1x128x128x128 tanh >> $O 2>&1
echo "--- racecheck" >> $O
timeout 400 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/diag_skel.py 2 1x20x24x36 tanh 2>&1 | grep -v "^$" | tail -40 >> $O
cat $O
