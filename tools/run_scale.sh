#!/bin/bash
# One multi-GPU measurement session on a box with N GPUs (gpurun --gpus N -- bash tools/run_scale.sh N [what...]):
#   time   tools/mgpu_time.py        where a sharded sweep's time goes (d=2500)
#   bench  bench.py --gpus N         the headline line (d=2500, 8 octaves, 1000 sweeps) incl. mgpu_check
#   d5000  bench.py --gpus N --division 5000 --octaves 12 --noise-dim 4    BASELINE configs[4]
# Results go to gpurun_out/ (copied to profiles/ by hand).
N=${1:-8}; shift
WHAT=${@:-time bench d5000}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for w in $WHAT; do
  case $w in
    time)  timeout 600 $TR --master-port 29511 tools/mgpu_time.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|^$" | tee gpurun_out/mgpu_time_n$N.log ;;
    bench) timeout 900 $TR --master-port 29512 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; head -c 1200 gpurun_out/bench_n$N.json; echo; tail -3 gpurun_out/bench_n$N.err ;;
    d5000) timeout 900 $TR --master-port 29513 bench.py --gpus $N --division 5000 --octaves 12 --noise-dim 4 > gpurun_out/bench_d5000_n$N.json 2> gpurun_out/bench_d5000_n$N.err; echo "d5000 rc=$?"; head -c 1200 gpurun_out/bench_d5000_n$N.json; echo; tail -3 gpurun_out/bench_d5000_n$N.err ;;
  esac
done
