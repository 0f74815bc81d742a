#!/bin/bash
# first GPU contact: layer tests (UMMA last, under a timeout), then module parity, then a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_layers.py -q -m gpu -k "not umma" -x --tb=short > gpurun_out/t_layers_simt.log 2>&1; echo "layers_simt rc=$?"
timeout 600 python -m pytest tests/test_gpu_layers.py -q -m gpu -k "umma" --tb=short > gpurun_out/t_layers_umma.log 2>&1; echo "layers_umma rc=$?"
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"
tail -5 gpurun_out/t_layers_simt.log gpurun_out/t_layers_umma.log gpurun_out/t_parity.log
