#!/bin/bash
# A/B of the gradient all-reduce overlap on N GPUs (run with `gpurun --gpus N`): bucketed overlap vs one all-reduce after the backward.
TAG=${2:-rXX}; N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
run() {  # name, extra env
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $3 \
      bench.py --gpus $N --steps 60 --warmup 9 --no-roofline --no-cpu-baseline > $OUT/${TAG}_bench_n${N}_$1.json 2> $OUT/${TAG}_bench_n${N}_$1.err
  echo "$1 exit $?"; python -c "import json,sys; d=json.load(open('$OUT/${TAG}_bench_n${N}_$1.json')); print('$1', d['n_gpus'], round(d['ms_per_step'],4), round(d['value']))"
}
timeout 400 python bench.py --gpus 1 --steps 60 --warmup 9 --no-roofline --no-cpu-baseline > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
python -c "import json; d=json.load(open('$OUT/${TAG}_bench_n1.json')); print('n1', round(d['ms_per_step'],4), round(d['value']))"
run overlap "HG_X=0" 29521
run flat "HG_NO_GRAD_OVERLAP=1" 29522
run overlap2 "HG_X=0" 29523
run flat2 "HG_NO_GRAD_OVERLAP=1" 29524
