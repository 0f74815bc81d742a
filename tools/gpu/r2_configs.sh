#!/bin/bash
# BASELINE configs 3/4/5 on N GPUs of one box (strong scaling over a fixed total) + the headline line at the same N.
#   gpurun --gpus N --timeout 1500 -- 'bash tools/gpu/r2_configs.sh N "3 4 5"'
N=${1:-1}; CFGS=${2:-"3"}
mkdir -p gpurun_out
run() { if [ "$N" = "1" ]; then python bench.py --gpus 1 "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; fi; }
for c in $CFGS; do
  timeout 900 bash -c "$(declare -f run); N=$N; run --config $c --steps 2 --warmup 1" > gpurun_out/config${c}_n${N}.jsonl 2> gpurun_out/config${c}_n${N}.err; echo "config $c N=$N rc=$?"; grep '^{' gpurun_out/config${c}_n${N}.jsonl | cut -c1-260
done
timeout 600 bash -c "$(declare -f run); N=$N; run --steps 5 --warmup 3 --no-cpu-baseline" > gpurun_out/headline_n${N}.json 2> gpurun_out/headline_n${N}.err; echo "headline N=$N rc=$?"; grep '^{' gpurun_out/headline_n${N}.json | cut -c1-420
