#!/bin/bash
# Round-2 evidence passes (one GPU each; one profiler invocation per gpurun call):
#   gpurun --timeout 3000 -- 'bash tools/gpu/r2_round_end.sh tests'    full GPU test suite, smoke, headline bench + reference arm, configs 1-2, parity margins
#   gpurun --timeout 1500 -- 'bash tools/gpu/r2_round_end.sh list'     ncu launch list of one steady-state step
#   gpurun --timeout 1500 -- 'bash tools/gpu/r2_round_end.sh layers'   per-launch metrics of one denoiser call
#   gpurun --timeout 1500 -- 'bash tools/gpu/r2_round_end.sh solver'   stamps/s of the classical solvers alone + one --set full capture of k_solver (Wiener)
#   gpurun --timeout 1500 -- 'bash tools/gpu/r2_round_end.sh full'     one --set full capture (source-level) of the four chain-kernel launches
mkdir -p gpurun_out
BCMD="python bench.py --steps 1 --warmup 3 --stamps 5000 --no-cpu-baseline"
case "$1" in
tests)
  timeout 1800 python -m pytest tests -q -m gpu --tb=short > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/t_gpu.log
  timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke.log
  timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench.json
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_ref.json
  timeout 600 python tools/bench_configs.py > gpurun_out/configs12.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"; cut -c1-220 gpurun_out/configs12.jsonl
  timeout 600 python tools/parity_report.py 48 > gpurun_out/parity.md 2> gpurun_out/parity.err; echo "parity rc=$?"
  ;;
list)
  timeout 600 $BCMD > gpurun_out/plain.log 2>&1 && \
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 200 --csv --log-file gpurun_out/launches.csv $BCMD > gpurun_out/ncu1.log 2>&1; echo "ncu list rc=$?"
  ;;
layers)
  timeout 600 $BCMD > gpurun_out/plain2.log 2>&1 && \
  timeout 1200 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"k_conv_umma|k_l1_chain|k_l2_chain|k_rb_umma|k_head|k_tail|k_g_xupdate|k_subnet|k_g_prologue|k_moments" -s 400 -c 30 --csv --log-file gpurun_out/layers.csv $BCMD > gpurun_out/ncu2.log 2>&1; echo "ncu layers rc=$?"
  ;;
full)
  timeout 600 $BCMD > gpurun_out/plain3.log 2>&1 && \
  timeout 1200 ncu --set full --import-source on --clock-control none -k regex:"k_l1_chain|k_l2_chain" -s 8 -c 4 -o gpurun_out/full_chain -f $BCMD > gpurun_out/ncu3.log 2>&1; echo "ncu full rc=$?"
  ;;
solver)
  timeout 600 python tools/gpu/solver_probe.py > gpurun_out/solver_probe.jsonl 2> gpurun_out/solver_probe.err; echo "probe rc=$?"; cat gpurun_out/solver_probe.jsonl
  [ -s gpurun_out/solver_probe.jsonl ] && timeout 900 ncu --set full --import-source on --clock-control none -k regex:"${SOLVER_KERNEL:-k_wiener48}" -s 1 -c 1 -o gpurun_out/full_solver -f python tools/gpu/solver_probe.py > gpurun_out/ncu4.log 2>&1; echo "ncu solver rc=$?"
  ;;
esac
