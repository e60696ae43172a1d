#!/bin/bash
# A/B of the higher-priority capture stream on N GPUs (run with `gpurun --gpus N`).
TAG=${2:-rXX}; N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
run() {  # name, extra env, port
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $3 \
      bench.py --gpus $N --steps 60 --warmup 9 --no-roofline --no-cpu-baseline 2> $OUT/${TAG}_bench_n${N}_$1.err | grep "^{" > $OUT/${TAG}_bench_n${N}_$1.json
  python -c "import json,sys; d=json.load(open('$OUT/${TAG}_bench_n${N}_$1.json')); print('$1', d['n_gpus'], round(d['ms_per_step'],4), round(d['value']))"
}
run prio_on "HG_CAPTURE_PRIORITY=1" 29521
run prio_off "HG_CAPTURE_PRIORITY=0" 29522
run prio_on2 "HG_CAPTURE_PRIORITY=1" 29523
run prio_off2 "HG_CAPTURE_PRIORITY=0" 29524
