#!/bin/bash
# source-level capture of two epilogue-bound launches of the third denoiser call: level-2 conv2 with U-Net skip, then up1
mkdir -p gpurun_out
BCMD="python bench.py --steps 1 --warmup 3 --stamps 5000 --no-cpu-baseline"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_conv_umma|k_rb_umma" -s 91 -c 2 -o gpurun_out/full_epi -f $BCMD > gpurun_out/ncu_full_epi.log 2>&1; echo "ncu full rc=$?"
