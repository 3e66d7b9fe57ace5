#!/bin/bash
# round 2, GPU call 17: bench with the IN fold; fused stride-2 dgrad for the wide k4 layers A/B
O=gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sliding > $O/r2c17_bench.json 2> $O/r2c17_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('$O/r2c17_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['roofline']['frac']);print(d['roofline']['families_ms_per_step'])"
for k in 32 64 128; do echo "K4MAX $k"; VG_TC_S2FUSED_K4MAX=$k timeout 100 python scripts/bench_conv.py dgrad 64-128 2>&1 | tail -1; VG_TC_S2FUSED_K4MAX=$k timeout 100 python scripts/bench_conv.py dgrad 128-256 2>&1 | tail -1; done
VG_TC_S2FUSED_K4MAX=128 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv" 2>&1 | tail -2
