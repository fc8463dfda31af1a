#!/bin/bash
# Round evidence on one B200 (under gpurun): full GPU test suite, smoke, both bench arms, ncu set.
TAG=${1:-r01v2}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 600 python bench.py ) > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 600 gpurun_out/bench_$TAG.err
( time timeout 600 python bench.py --impl reference ) > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
tail -c 300 gpurun_out/bench_ref_$TAG.err
timeout 900 bash profiles/run_ncu_final.sh $TAG 2>&1 | tail -8
