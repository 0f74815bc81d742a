#!/bin/bash
# tests -> smoke -> bench -> ncu launch list -> ncu full capture of the tcgen05 conv kernel
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --tb=short > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/t_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
BCMD="python bench.py --steps 1 --warmup 3 --stamps 1024 --no-cpu-baseline"
timeout 600 $BCMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 300 --csv --log-file gpurun_out/launches.csv $BCMD > gpurun_out/ncu1.log 2>&1; echo "ncu list rc=$?"
timeout 600 $BCMD > gpurun_out/plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_conv_umma -s 340 -c 34 -o gpurun_out/prof_umma $BCMD > gpurun_out/ncu2.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
