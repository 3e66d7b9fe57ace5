#!/bin/bash
# round 2, GPU call 18: full GPU suite + bench after the IN fold (generic + forward halo pass) and the fused k4 s2 dgrad default
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r2c18_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -5 $O/r2c18_pytest_all.log
timeout 100 python scripts/bench_in.py 2>&1 | grep "^apply\|^bwd" | tee $O/r2c18_in.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sliding > $O/r2c18_bench.json 2> $O/r2c18_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('$O/r2c18_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['roofline']['frac']);print(d['roofline']['families_ms_per_step'])"
