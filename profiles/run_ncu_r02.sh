#!/bin/bash
# Evidence set of round 2 (run under gpurun, one GPU): launch list of one bench step, full captures of K1
# (roofline size, 2 M reads), the K4 DP (throughput shape, full grid) and the K5 layer kernel (one launch of
# the draft step at depth ~150 of 20 clusters).
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-consensus --no-modules --no-concurrent --no-stream --roofline-reads 1000000 > $OUT/launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k1_stream -s 6 -c 1 -f -o $OUT/prof_k1stream_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-consensus --no-modules --no-concurrent --no-stream --reads 100000 --roofline-reads 2000000 > $OUT/prof_k1stream_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k4t_dp -s 1 -c 1 -f -o $OUT/prof_k4tdp_$TAG \
    python scripts/k4_probe.py 12000 0 > $OUT/prof_k4tdp_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k5r_layer -s 150 -c 1 -f -o $OUT/prof_k5r_$TAG \
    python scripts/poa_depth_probe.py 40000 200 > $OUT/prof_k5r_$TAG.log 2>&1
# map kernel (hash-table probing) and K0 (quality statistics): the launches of one bench step; summarize.py lists every
# captured launch, the largest grid is the last tile's (added at the end of round 2, not run: the GPU budget was spent)
ncu --set full --clock-control none --import-source on -k regex:k2_map_kernel -s 8 -c 12 -f -o $OUT/prof_k2map_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-consensus --no-modules --no-roofline --no-concurrent --no-stream > $OUT/prof_k2map_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k0_quality -s 2 -c 1 -f -o $OUT/prof_k0_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-consensus --no-modules --no-roofline --no-concurrent --no-stream > $OUT/prof_k0_$TAG.log 2>&1
ls -la $OUT | tail -8
