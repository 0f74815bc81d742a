#!/bin/bash
mkdir -p gpurun_out
export GDECONV_CHAIN=1
BCMD="python bench.py --steps 1 --warmup 3 --stamps 5000 --no-cpu-baseline"
timeout 600 $BCMD > gpurun_out/plain2.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct --clock-control none -k regex:"k_conv" -s 120 -c 26 --csv --log-file gpurun_out/layers_chain.csv $BCMD > gpurun_out/ncu2.log 2>&1; echo "ncu rc=$?"
for DJ in 1 2 6; do
GDECONV_CHAIN_DJ=$DJ timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dj$DJ.json 2>> gpurun_out/bench.err; python -c "
import json;d=json.load(open('gpurun_out/bench_dj$DJ.json'));print('dj $DJ', d['value'], d['ms_per_step'])"
done
