timeout 400 python -m pytest tests -m gpu -x -q -k "clustering or golden or many_representatives or rank_sharded" 2>&1 | tail -4
for v in "A=1" "NGSID_NO_PREFETCH=1" "NGSID_MAP_BPS=4" "NGSID_MAP_BPS=3"; do
  env $v timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu --no-consensus --no-roofline > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err
  python - "$v" <<'P'
import json,sys
d=json.loads([l for l in open("gpurun_out/bench_p.json") if l.startswith("{")][-1])
print(sys.argv[1], round(d["value"]), round(d["e2e"]["value"]), d["phase_ms_per_step"], {k:d["cluster_stats"][k] for k in ("n_alignments","n_chain_steps","n_tiles","n_new_reps","n_aln_passed","n_mapped")}, d["gpu_launches"])
P
done
