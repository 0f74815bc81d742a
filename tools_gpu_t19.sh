#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/t_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['gpu_launches'], d['e2e']['value'], d['clocks'])"
timeout 900 python tools/bench_configs.py --solver-stamps 200000 --sweep-stamps 256 2>&1 | grep '"config": 4\|"config": 1'
BCMD="python bench.py --steps 1 --warmup 3 --stamps 5000 --no-cpu-baseline"
timeout 600 $BCMD > gpurun_out/plain2.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"k_head|k_tail|k_g_xupdate|k_subnet|k_g_prologue|k_moments" -c 12 --csv --log-file gpurun_out/layers_small.csv $BCMD > gpurun_out/ncu2.log 2>&1; echo "ncu rc=$?"
