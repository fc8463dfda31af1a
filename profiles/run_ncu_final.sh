#!/bin/bash
# Evidence set of a round (run under gpurun, one GPU): launch list of one bench step, full captures
# of K1 (roofline size), the K4 DP (throughput shape, full grid) and traceback, and the K5 wavefront kernel.
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-consensus --roofline-reads 1000000 > $OUT/launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k1_stream -s 6 -c 1 -f -o $OUT/prof_k1stream_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-consensus --reads 100000 --roofline-reads 2000000 > $OUT/prof_k1stream_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k4t_dp -s 1 -c 1 -f -o $OUT/prof_k4tdp_$TAG \
    python scripts/k4_probe.py 12000 0 > $OUT/prof_k4tdp_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k4t_traceback -s 1 -c 1 -f -o $OUT/prof_k4ttb_$TAG \
    python scripts/k4_probe.py 12000 0 > $OUT/prof_k4ttb_$TAG.log 2>&1
ncu --set full --clock-control none -k regex:k5w_poa -c 1 -f -o $OUT/prof_k5w_$TAG \
    python scripts/poa_depth_probe.py 40000 30 > $OUT/prof_k5w_$TAG.log 2>&1
ls -la $OUT | tail -12
