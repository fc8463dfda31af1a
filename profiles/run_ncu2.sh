#!/bin/bash
# K1 fast kernel at roofline size + three large K4 launches (tile 65536)
set -u
TAG=${1:-r01b}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu --reads 100000 --roofline-reads 2000000 --tile 65536"
ncu --set full --clock-control none --import-source on -k regex:k1_fast -s 6 -c 1 -f -o $OUT/prof_k1fast_$TAG $BENCH > $OUT/prof_k1fast_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k4_align -s 18 -c 3 -f -o $OUT/prof_k4_$TAG $BENCH > $OUT/prof_k4_$TAG.log 2>&1
ls -la $OUT
