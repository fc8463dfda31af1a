#!/bin/bash
# Reproduces the ncu captures kept under profiles/ (run under gpurun, one GPU).
#   profiles/run_ncu.sh <round-tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu --reads 100000 --roofline-reads 1000000"
# 1. every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/launches_$TAG.log 2>&1
# 2. K1 at the roofline size (launches 0-3 are the 100k-read steps, later ones the replicated input)
ncu --set full --clock-control none --import-source on -k regex:k1_minimizers -s 6 -c 1 -f -o $OUT/prof_k1_$TAG $BENCH > $OUT/prof_k1_$TAG.log 2>&1
# 3. K4 (largest launch of a step)
ncu --set full --clock-control none --import-source on -k regex:k4_align -s 12 -c 1 -f -o $OUT/prof_k4_$TAG $BENCH > $OUT/prof_k4_$TAG.log 2>&1
# 4. the map kernel
ncu --set full --clock-control none --import-source on -k regex:k2_map -s 20 -c 1 -f -o $OUT/prof_map_$TAG $BENCH > $OUT/prof_map_$TAG.log 2>&1
ls -la $OUT
