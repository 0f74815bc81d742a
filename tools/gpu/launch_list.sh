#!/bin/bash
# ncu launch list (gpu__time_duration only) of ONE steady-state step of the headline command (2 chunks of 5000 stamps:
# SubNet, prologue, 8 x [x-update, head, 36 tcgen05 conv launches, tail gather], moments), after the command ran clean without ncu.
mkdir -p gpurun_out
BCMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 $BCMD > gpurun_out/plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 1120 -c 640 --csv --log-file gpurun_out/launches.csv $BCMD > gpurun_out/ncu1.log 2>&1; echo "ncu list rc=$?"
