#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -x > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"; tail -n 8 gpurun_out/t_parity.log
for CH in 0 1; do
GDECONV_CHAIN=$CH timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_chain$CH.json 2>> gpurun_out/bench.err; python -c "
import json;d=json.load(open('gpurun_out/bench_chain$CH.json'));print('chain $CH', d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value'], d['clocks'])"
done
tail -n 5 gpurun_out/bench.err
