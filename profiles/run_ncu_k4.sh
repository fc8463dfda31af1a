#!/bin/bash
# K4 trace variant: one large DP launch + its traceback (tile 65536)
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k4t_ -s 24 -c 2 -f -o gpurun_out/prof_k4t_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-consensus --no-roofline --reads 100000 > gpurun_out/prof_k4t_$TAG.log 2>&1
