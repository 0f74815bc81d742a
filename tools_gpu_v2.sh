#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_layers.py -q -m gpu --tb=short -x > gpurun_out/t_layers.log 2>&1; echo "layers rc=$?"; tail -n 3 gpurun_out/t_layers.log
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -q -m gpu --tb=short > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"; tail -n 5 gpurun_out/t_parity.log
for SUB in 64 128 256 512; do
  GDECONV_SUBCHUNK=$SUB timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sub$SUB.json 2>> gpurun_out/bench.err; echo "sub=$SUB"; python -c "
import json;d=json.load(open('gpurun_out/bench_sub$SUB.json'));print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['gpu_launches'])"
done
BCMD="python bench.py --steps 1 --warmup 3 --stamps 1024 --no-cpu-baseline"
export GDECONV_SUBCHUNK=512
timeout 600 $BCMD > gpurun_out/plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_conv_umma -s 340 -c 34 -o gpurun_out/prof_umma_v2 $BCMD > gpurun_out/ncu2.log 2>&1; echo "ncu full rc=$?"
