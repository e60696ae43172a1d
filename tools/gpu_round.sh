#!/bin/bash
# One gpurun call = one measurement round.  Usage (from the repo root, on the GPU box):
#   bash tools/gpu_round.sh TAG [tests] [bench] [micro] [launches] [step] [ncu:<targets,comma-separated>]
# With no stage names every stage runs.  Everything lands under gpurun_out/TAG_* (merged back by gpurun).
set -u
TAG=${1:-rXX}; shift || true
STAGES="$*"
[ -z "$STAGES" ] && STAGES="tests bench micro launches step ncu:rotate,rotate_cl,adain_cl,final_conv,conv"
OUT=gpurun_out
mkdir -p $OUT
has() { case " $STAGES " in *" $1 "*) return 0;; esac; return 1; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1

if has tests; then
  timeout 600 python -m pytest tests -m gpu ${PYTEST_X--x} -q > $OUT/${TAG}_pytest_gpu.txt 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.txt
  tail -5 $OUT/${TAG}_pytest_gpu.txt
fi
if has bench; then
  timeout 900 python bench.py --steps 60 --warmup 9 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
  echo "bench exit $?"; cat $OUT/${TAG}_bench.json
fi
if has bench128; then
  timeout 600 python bench.py --steps 30 --warmup 6 --img-size 128 --batch 32 --no-cpu-baseline --no-roofline > $OUT/${TAG}_bench128.json 2> $OUT/${TAG}_bench128.err
  echo "bench128 exit $?"; cat $OUT/${TAG}_bench128.json
fi
if has sweep; then
  timeout 600 python tools/sweep_bench.py --img-size 64 > $OUT/${TAG}_sweep.txt 2>&1
  timeout 600 python tools/sweep_bench.py --img-size 128 >> $OUT/${TAG}_sweep.txt 2>&1
  echo "sweep exit $?"; grep metric $OUT/${TAG}_sweep.txt
fi
if has tests_cl; then
  timeout 600 python -m pytest tests/test_gpu_channels_last.py tests/test_gpu_generator.py -m gpu -q > $OUT/${TAG}_pytest_cl.txt 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest_cl.txt
  tail -3 $OUT/${TAG}_pytest_cl.txt
fi
if has pipe; then
  timeout 600 python tools/microbench.py pipeline > $OUT/${TAG}_microbench_pipeline.txt 2>&1
  echo "microbench pipeline exit $?"; grep -E "adain_cl block|spectral|rotate_cl|final_conv" $OUT/${TAG}_microbench_pipeline.txt
fi
if has mma; then
  # tensor-core final-layer kernels forced on (HG_FINAL_CONV_MMA=7): pipeline microbench + a short bench
  HG_FINAL_CONV_MMA=7 timeout 300 python tools/microbench.py pipeline > $OUT/${TAG}_microbench_pipeline_mma.txt 2>&1
  grep -E "final_conv" $OUT/${TAG}_microbench_pipeline_mma.txt
  HG_FINAL_CONV_MMA=7 timeout 300 python bench.py --steps 30 --warmup 6 --no-cpu-baseline --no-roofline > $OUT/${TAG}_bench_mma.json 2> $OUT/${TAG}_bench_mma.err
  echo "bench mma exit $?"; cut -c1-220 $OUT/${TAG}_bench_mma.json
fi
if has optin; then
  # code finished after round 1's GPU budget was spent: first measurements
  timeout 900 python -m pytest tests -m gpu -q --runxfail -k "slab32 or k5_in_child or tcgen05_convs or wgrad_grouped" > $OUT/${TAG}_pytest_optin.txt 2>&1
  echo "optin pytest exit $?"; tail -15 $OUT/${TAG}_pytest_optin.txt
  HG_BENCH_SLAB32=1 timeout 600 python tools/microbench.py rotate32 > $OUT/${TAG}_microbench_rotate32.txt 2>&1
  grep -E "32\^3" $OUT/${TAG}_microbench_rotate32.txt
  HG_WGRAD_GROUP=1 timeout 300 python tools/microbench.py conv > $OUT/${TAG}_microbench_conv_grouped.txt 2>&1
  grep -E "^block|^proj|totals" $OUT/${TAG}_microbench_conv_grouped.txt
  HG_WGRAD_GROUP=1 timeout 300 python bench.py --steps 30 --warmup 6 --no-cpu-baseline --no-roofline > $OUT/${TAG}_bench_wgrad_group.json 2> $OUT/${TAG}_bench_wgrad_group.err
  echo "bench wgrad_group exit $?"; cut -c1-220 $OUT/${TAG}_bench_wgrad_group.json
  HG_D_TCGEN05=1 timeout 300 python bench.py --steps 30 --warmup 6 --no-cpu-baseline --no-roofline > $OUT/${TAG}_bench_d_tcgen05.json 2> $OUT/${TAG}_bench_d_tcgen05.err
  echo "bench d_tcgen05 exit $?"; cut -c1-220 $OUT/${TAG}_bench_d_tcgen05.json; tail -3 $OUT/${TAG}_bench_d_tcgen05.err
fi
if has micro; then
  timeout 900 python tools/microbench.py all > $OUT/${TAG}_microbench.txt 2>&1
  echo "microbench exit $?"
fi
if has rot16; then
  timeout 600 python tools/microbench.py rotate16 > $OUT/${TAG}_rotate16.txt 2>&1
  echo "rotate16 exit $?"
fi
if has step; then
  timeout 600 python tools/profile_step.py 6 > $OUT/${TAG}_torch_profiler_step.txt 2>&1
  echo "profile_step exit $?"
fi
if has launches; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-roofline > $OUT/${TAG}_launches_run.log 2>&1
  echo "launch list exit $?"
fi
for st in $STAGES; do
  case $st in ncu:*)
    targets=$(echo ${st#ncu:} | tr ',' ' ')
    for t in $targets; do
      HG_NCU_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:hg:: -c ${NCU_COUNT:-40} -f \
          -o $OUT/${TAG}_full_$t python tools/ncu_targets.py $t > $OUT/${TAG}_full_$t.log 2>&1
      echo "ncu full $t exit $?"
      ncu -i $OUT/${TAG}_full_$t.ncu-rep --page raw --csv > $OUT/${TAG}_full_${t}_raw.csv 2>/dev/null
      # gpurun merges at most 64 MiB back: keep the report only when it is small, the raw CSV always
      [ $(stat -c %s $OUT/${TAG}_full_$t.ncu-rep) -gt 12000000 ] && rm -f $OUT/${TAG}_full_$t.ncu-rep
    done;;
  esac
done
ls -la $OUT | head -50
