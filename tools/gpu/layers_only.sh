#!/bin/bash
# per-launch table of one denoiser call with the tensor-pipe OCCUPANCY (incl. shared-memory operand fetch) next to the math activity
mkdir -p gpurun_out
BCMD="python bench.py --steps 1 --warmup 3 --stamps 5000 --no-cpu-baseline"
timeout 1200 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"k_conv_umma|k_rb_umma|k_head|k_tail|k_g_xupdate" -s 185 -c 40 --csv --log-file gpurun_out/layers_tc.csv $BCMD > gpurun_out/ncu_tc.log 2>&1; echo "ncu layers rc=$?"
