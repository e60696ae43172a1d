#!/bin/bash
# Multi-GPU measurements on N GPUs (run with `gpurun --gpus N`): BASELINE cfg 2 (64x64, B=64), cfg 4 (128x128, B=32) and cfg 5 (sweep).
TAG=${2:-rXX}; N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
one() {  # name, port, bench args...
  name=$1; port=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@" \
      2> $OUT/${TAG}_${name}_n$N.err | grep "^{" > $OUT/${TAG}_${name}_n$N.json
  python -c "import json; d=json.load(open('$OUT/${TAG}_${name}_n$N.json')); print('$name n=$N', round(d['ms_per_step'],4), 'ms/step', round(d['value']), 'img/s  e2e', round(d['e2e']['value']))" || tail -5 $OUT/${TAG}_${name}_n$N.err
}
s1() {  # single-GPU reference on the same box
  name=$1; shift
  timeout 600 python bench.py --gpus 1 "$@" 2> $OUT/${TAG}_${name}_n1.err | grep "^{" > $OUT/${TAG}_${name}_n1.json
  python -c "import json; d=json.load(open('$OUT/${TAG}_${name}_n1.json')); print('$name n=1', round(d['ms_per_step'],4), 'ms/step', round(d['value']), 'img/s  e2e', round(d['e2e']['value']))"
}
s1 train64 --steps 60 --warmup 9 --no-roofline --no-cpu-baseline
one train64 29531 --steps 60 --warmup 9 --no-roofline --no-cpu-baseline
s1 train128 --steps 60 --warmup 9 --img-size 128 --batch 32 --no-roofline --no-cpu-baseline
one train128 29532 --steps 60 --warmup 9 --img-size 128 --batch 32 --no-roofline --no-cpu-baseline
if [ "${3:-}" = "sweep" ]; then
  s1 sweep128 --mode sweep --steps 3 --warmup 3
  one sweep128 29533 --mode sweep --steps 3 --warmup 3
fi
