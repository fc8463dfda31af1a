#!/bin/bash
# K5 wavefront kernel, draft of 20 clusters x 50 reads: full capture with source counters
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k5w_poa -c 1 -f -o gpurun_out/prof_k5w_$TAG \
    python scripts/poa_depth_probe.py 40000 30 > gpurun_out/prof_k5w_$TAG.log 2>&1
tail -2 gpurun_out/prof_k5w_$TAG.log
