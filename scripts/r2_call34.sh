#!/bin/bash
# round 2, GPU call 34: InstanceNorm statistics re-used (cached per tensor, derived for upsample+concat): parity + bench A/B
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r2c34_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2c34_pytest.log
for sr in 1 0; do
  VG_STATS_REUSE=$sr timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sliding > $O/r2c34_bench_reuse$sr.json 2>/dev/null; echo "bench reuse=$sr rc=$?"; python - <<PY
import json
d=json.loads(open('$O/r2c34_bench_reuse$sr.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value']); f=d['roofline']['families_ms_per_step']; print({k:f[k] for k in f if 'instnorm' in k})
PY
done
