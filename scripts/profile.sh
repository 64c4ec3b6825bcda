#!/bin/bash
# ncu evidence for the bench command (run under gpurun, 1 GPU): launch list + one full capture of the
# element kernel.  Usage: scripts/profile.sh <tag> [extra bench args]
tag=${1:-r1}; shift
mkdir -p gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --pcg-iters 3 $*"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $BENCH > gpurun_out/launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'hex_fast|hex_apply|hex_patch' -s 3 -c 2 -f -o gpurun_out/prof_$tag $BENCH > gpurun_out/prof_$tag.log 2>&1
ls -la gpurun_out | tail -8
