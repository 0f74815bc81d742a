#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -x > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"; tail -n 5 gpurun_out/t_parity.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2>> gpurun_out/bench.err; python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print(d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['clocks'])"
BCMD="python bench.py --steps 1 --warmup 3 --stamps 5000 --no-cpu-baseline"
timeout 600 $BCMD > gpurun_out/plain2.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"k_head|k_tail|k_g_xupdate" -c 6 --csv --log-file gpurun_out/layers_small.csv $BCMD > gpurun_out/ncu2.log 2>&1; echo "ncu rc=$?"
