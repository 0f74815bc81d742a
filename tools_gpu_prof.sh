#!/bin/bash
mkdir -p gpurun_out
BCMD="python bench.py --steps 1 --warmup 3 --stamps 1024 --no-cpu-baseline"
export GDECONV_SUBCHUNK=512
timeout 600 $BCMD > gpurun_out/plain2.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_conv_umma -s 340 -c 19 -o gpurun_out/prof_umma_v4 $BCMD > gpurun_out/ncu2.log 2>&1; echo "ncu full rc=$?"
