#!/bin/bash
mkdir -p gpurun_out
for CH in 4096 5000 10000; do
GDECONV_CHUNK=$CH GDECONV_SUBCHUNK=100000 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c$CH.json 2>> gpurun_out/bench.err; python -c "
import json;d=json.load(open('gpurun_out/bench_c$CH.json'));print('chunk $CH', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['gpu_launches'], d['e2e']['value'])"
done
export GDECONV_CHUNK=5000 GDECONV_SUBCHUNK=100000
BCMD="python bench.py --steps 1 --warmup 3 --stamps 5000 --no-cpu-baseline"
timeout 600 $BCMD > gpurun_out/plain2.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"k_conv_umma|k_head|k_tail|k_g_xupdate|k_subnet|k_g_prologue|k_moments" -s 185 -c 40 --csv --log-file gpurun_out/layers_big.csv $BCMD > gpurun_out/ncu2.log 2>&1; echo "ncu rc=$?"
