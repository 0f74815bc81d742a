#!/bin/bash
mkdir -p gpurun_out
export GDECONV_CHUNK=2048 GDECONV_SUBCHUNK=100000
BCMD="python bench.py --steps 1 --warmup 3 --stamps 2048 --no-cpu-baseline"
timeout 600 $BCMD > gpurun_out/plain2.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_conv_umma|k_head|k_tail" -s 174 -c 14 -o gpurun_out/prof_mem_v5 $BCMD > gpurun_out/ncu2.log 2>&1; echo "ncu full rc=$?"
