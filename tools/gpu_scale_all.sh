#!/bin/bash
# Weak-scaling table from ONE 8-GPU box (run with `gpurun --gpus 8`): BASELINE cfg 2 at N = 1, 2, 4, 8 back to back, then
# cfg 4 (128x128, B=32) and cfg 5 (view sweep) at N = 1 and N = 8.  Usage: bash tools/gpu_scale_all.sh TAG
TAG=${1:-rXX}; OUT=gpurun_out; mkdir -p $OUT
run() {  # name, N, port, bench args...
  name=$1; N=$2; port=$3; shift 3
  if [ "$N" = 1 ]; then
    timeout 600 python bench.py --gpus 1 "$@" 2> $OUT/${TAG}_${name}_n1.err | grep "^{" > $OUT/${TAG}_${name}_n1.json
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@" \
        2> $OUT/${TAG}_${name}_n$N.err | grep "^{" > $OUT/${TAG}_${name}_n$N.json
  fi
  python -c "import json; d=json.load(open('$OUT/${TAG}_${name}_n$N.json')); print('$name n=$N', round(d['ms_per_step'],4), 'ms/step', round(d['value']), 'img/s  e2e', round(d['e2e']['value']))" || tail -5 $OUT/${TAG}_${name}_n$N.err
}
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/${TAG}_gpus.txt 2>&1
A="--steps 60 --warmup 9 --no-roofline --no-cpu-baseline"
run train64 1 0 $A
run train64 2 29541 $A
run train64 4 29542 $A
run train64 8 29543 $A
run train128 1 0 $A --img-size 128 --batch 32
run train128 8 29544 $A --img-size 128 --batch 32
run sweep128 1 0 --mode sweep --steps 3 --warmup 3
run sweep128 8 29545 --mode sweep --steps 3 --warmup 3
