#!/bin/bash
# One GPU call of a build->measure cycle (run under gpurun from the repo root):
#   profiles/gpu_round.sh <tag> [tests] [bench] [ncu] [probe]
# Writes everything under gpurun_out/<tag>_*; copy the digests you want judged into profiles/.
tag=$1; shift
what="$*"; [ -z "$what" ] && what="tests bench ncu"
mkdir -p gpurun_out
SMALL="python bench.py --nseq ${NSEQ:-37888} --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
for w in $what; do
  case $w in
    tests) python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log;;
    bench) python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json;;
    probe) python profiles/probe_fp64.py > gpurun_out/${tag}_probe.txt 2>&1; cat gpurun_out/${tag}_probe.txt;;
    ncu)
      $SMALL > gpurun_out/${tag}_plain_small.log 2>&1 &&
      ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $SMALL > gpurun_out/${tag}_ncu_launches.log 2>&1
      echo "launch list rc=$?"
      $SMALL > gpurun_out/${tag}_plain_small2.log 2>&1 &&
      ncu --set full --clock-control none --import-source on -k regex:k_mma_estep -c 1 -f -o gpurun_out/${tag}_prof $SMALL > gpurun_out/${tag}_ncu_full.log 2>&1
      echo "ncu full rc=$?";;
  esac
done
