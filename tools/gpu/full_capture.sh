#!/bin/bash
# one `ncu --set full` capture of the tcgen05 conv launches of (most of) one denoiser call, after the command ran clean without ncu
mkdir -p gpurun_out
BCMD="python bench.py --steps 1 --warmup 3 --stamps 5000 --no-cpu-baseline"
timeout 600 $BCMD > gpurun_out/plain_full.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none -k regex:"k_conv_umma|k_rb_umma" -s 196 -c 8 -o gpurun_out/full_v9 -f $BCMD > gpurun_out/ncu_full_v9.log 2>&1; echo "ncu full rc=$?"
