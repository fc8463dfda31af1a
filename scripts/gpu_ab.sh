#!/bin/bash
# A/B runs of the clustering step on one B200: each line = one environment setting
run() {
  env $1 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-consensus --no-roofline $BENCH_ARGS > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
  python - "$1" <<'P'
import json,sys
d=json.loads([l for l in open("gpurun_out/bench_ab.json") if l.startswith("{")][-1])
print(sys.argv[1], round(d["value"]), round(d["e2e"]["value"]), {k: round(v,2) for k,v in d["phase_ms_per_step"].items()}, {k:d["cluster_stats"][k] for k in ("n_alignments","n_chain_steps","n_tiles","n_map_launch_reads")}, d["gpu_launches"])
P
}
for v in "$@"; do run "$v"; done
