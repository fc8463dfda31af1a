#!/bin/bash
# K1 stream kernel at the roofline size (2 M reads per launch): full capture with source counters
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k1_stream -s 6 -c 1 -f -o gpurun_out/prof_k1stream_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-consensus --reads 100000 --roofline-reads 2000000 > gpurun_out/prof_k1stream_$TAG.log 2>&1
tail -3 gpurun_out/prof_k1stream_$TAG.log
