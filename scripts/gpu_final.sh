#!/bin/bash
# round-end validation: all GPU parity tests, smoke, both bench arms, the secondary configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 python scripts/bench_configs.py all > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"; cat gpurun_out/configs.jsonl; tail -3 gpurun_out/configs.err
