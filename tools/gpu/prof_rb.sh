#!/bin/bash
# full-section ncu capture of the four fused ResBlock launches of one denoiser call
mkdir -p gpurun_out
BCMD="python bench.py --steps 1 --warmup 3 --stamps 5000 --no-cpu-baseline"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_rb_umma" -s 4 -c 2 -o gpurun_out/full_rb -f $BCMD > gpurun_out/ncu_full_rb.log 2>&1; echo "ncu full rc=$?"
