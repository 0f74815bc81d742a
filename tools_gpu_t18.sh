#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/t_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json | cut -c1-1500
timeout 1500 python tools/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"; cat gpurun_out/configs.jsonl; tail -n 5 gpurun_out/configs.err
