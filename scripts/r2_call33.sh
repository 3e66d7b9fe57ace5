#!/bin/bash
# round 2, GPU call 33: conv bias gradients taken inside the consuming norm's backward (sinks): parity + bench A/B
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r2c33_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2c33_pytest.log
for bs in 1 0; do
  VG_BIAS_SINKS=$bs timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sliding > $O/r2c33_bench_sinks$bs.json 2>/dev/null; echo "bench sinks=$bs rc=$?"; python - <<PY
import json
d=json.loads(open('$O/r2c33_bench_sinks$bs.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac']); print(d['roofline']['families_ms_per_step'])
PY
done
