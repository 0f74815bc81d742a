#!/bin/bash
mkdir -p gpurun_out
for SUB in 64 128 256 512; do
  GDECONV_SUBCHUNK=$SUB timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sub$SUB.json 2>> gpurun_out/bench.err; echo "sub=$SUB"; python -c "
import json;d=json.load(open('gpurun_out/bench_sub$SUB.json'));print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['gpu_launches'])"
done
GDECONV_CHUNK=1024 GDECONV_SUBCHUNK=128 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c1024.json 2>> gpurun_out/bench.err; python -c "
import json;d=json.load(open('gpurun_out/bench_c1024.json'));print('chunk1024 sub128', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['gpu_launches'])"
GDECONV_CHUNK=2048 GDECONV_SUBCHUNK=2048 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2048.json 2>> gpurun_out/bench.err; python -c "
import json;d=json.load(open('gpurun_out/bench_c2048.json'));print('chunk2048 sub2048', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['gpu_launches'])"
