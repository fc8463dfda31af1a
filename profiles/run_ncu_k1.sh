#!/bin/bash
# K1 fast kernel at the roofline size (2 M reads per launch)
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k1_fast -s 6 -c 1 -f -o gpurun_out/prof_k1fast_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-consensus --reads 100000 --roofline-reads 2000000 > gpurun_out/prof_k1fast_$TAG.log 2>&1
