#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_layers.py -q -m gpu --tb=short -x > gpurun_out/t_layers.log 2>&1; echo "layers rc=$?"; tail -n 2 gpurun_out/t_layers.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"; tail -n 3 gpurun_out/t_parity.log
export GDECONV_CHUNK=5000 GDECONV_SUBCHUNK=100000
for BS in 6 8; do
GDECONV_BSTAGES=$BS timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bs$BS.json 2>> gpurun_out/bench.err; python -c "
import json;d=json.load(open('gpurun_out/bench_bs$BS.json'));print('bstages $BS', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['gpu_launches'], d['e2e']['value'])"
done
BCMD="python bench.py --steps 1 --warmup 3 --stamps 5000 --no-cpu-baseline"
timeout 600 $BCMD > gpurun_out/plain2.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"k_conv_umma|k_head|k_tail|k_g_xupdate|k_subnet|k_g_prologue|k_moments" -s 185 -c 40 --csv --log-file gpurun_out/layers_big.csv $BCMD > gpurun_out/ncu2.log 2>&1; echo "ncu rc=$?"
