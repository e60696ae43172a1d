#!/bin/bash
# Multi-GPU check of bench.py (run with `gpurun --gpus N`): N=1 and N=$1 back to back, as the driver launches them.
TAG=${2:-rXX}; N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_channels_last.py -m gpu -q 2>&1 | tail -2
timeout 400 python bench.py --gpus 1 --steps 30 --warmup 6 --no-roofline --no-cpu-baseline > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
echo "n1 exit $?"; cat $OUT/${TAG}_bench_n1.json | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 30 --warmup 6 --no-roofline > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
echo "n$N exit $?"; cat $OUT/${TAG}_bench_n$N.json | cut -c1-300; tail -5 $OUT/${TAG}_bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $OUT/${TAG}_ref_n$N.json 2> $OUT/${TAG}_ref_n$N.err
echo "ref n$N exit $?"; cat $OUT/${TAG}_ref_n$N.json | cut -c1-300
