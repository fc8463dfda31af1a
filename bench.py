#!/usr/bin/env python
"""
bench.py -- reads/s clustered + consensus bp/s on synthetic 750 bp ONT amplicon reads on N B200s
(BASELINE.json `metric`).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...       # the unmodified reference on the host cores

Workloads (config.workload names the one that ran):
  every N     : BASELINE.json configs[1] per GPU: 100 k reads, 10 species, k=13 w=20 (weak scaling:
                N x 100 k reads, `--t N` semantics of modules/parallelize.py);
                the consensus leg is configs[2] (cluster + POA consensus + 3 racon-style rounds).
  N = 8 also  : `north_star_configs3` = BASELINE.json configs[3] end to end: 10^6 reads, 50 species,
                cluster + consensus with --abundance_ratio 0.005 (the same JSON shape, nested).
  --config c1 / c3 forces one of them as the main line at any N (reads per GPU: 100 k / 125 k).

One "step" = one pass of the clustering hot path over the whole batch of every rank: K1 minimizers
+ K0 quality statistics + the greedy pass (K2/K3 mapping, K4 block alignment) and, for N > 1, the
NCCL gather of the surviving representatives + the log2(N) merge rounds.
  value   : reads/s with the reads resident in HBM, wall clock over all K steps between device syncs, max over
            ranks; the local pass of step s + 1 overlaps the exchange + merge rounds of step s
            (Pipeline.cluster_stream: second batch engine, second host thread, collectives on one thread)
  e2e     : the same from pinned host buffers (H2D of bases + qualities inside the timed region,
            D2H of the assignments), input double-buffered: the transfer of step n+1 runs on a second
            engine under the clustering pass of step n (Pipeline.prefetch)
  consensus: consensus bp/s of draft + reverse-complement merge + polishing of the FINAL clusters
  roofline: K1 (minimizer extraction) timed alone with CUDA events on a replicated input far larger
            than L2; algorithmic bytes = packed read + (offset, len) + 8 B / minimizer + count
  cpu_baseline: the unmodified reference (baseline/_ref) on a bounded sample, all host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K, W = 13, 20
CONFIGS = {
    "c1": {"reads_per_gpu": 100000, "species": 10, "abundance_ratio": 0.02, "max_seqs": 200, "racon_iter": 3,
           "name": "BASELINE.json configs[1] (cluster-only; consensus leg = configs[2])"},
    "c3": {"reads_per_gpu": 125000, "species": 50, "abundance_ratio": 0.005, "max_seqs": 200, "racon_iter": 3,
           "name": "BASELINE.json configs[3] (cluster + consensus)"},
    # configs[4] shape per GPU (5 M reads on 8 GPUs = 625 k per GPU; the default here is a bounded 200 k):
    # mixed-length 500-2000 bp PacBio-profile reads, --isoseq parameters k=15 w=50 (NGSpeciesID:264-269)
    "c4": {"reads_per_gpu": 200000, "species": 10, "abundance_ratio": 0.02, "max_seqs": 100, "racon_iter": 3,
           "k": 15, "w": 50, "profile": "pacbio", "len": (1900, 2000), "per_read_len": (500, 2000),
           "name": "BASELINE.json configs[4] shape (PacBio profile, k=15 w=50, cluster + consensus)"},
}


# ------------------------------------------------------------------------------------ workload
def _vector_scores_block(qual, offsets, k):
    p = np.minimum(10.0 ** (-(qual.astype(np.float64) - 33.0) / 10.0), 0.79433)
    lg = np.log1p(-p)
    cs = np.concatenate([[0.0], np.cumsum(lg)])
    n = len(offsets) - 1
    win = cs[k:] - cs[:-k]                     # window starting at flat position i
    lens = np.diff(offsets)
    valid = np.zeros(len(qual), dtype=bool)
    idx = np.arange(len(qual))
    read_of = np.repeat(np.arange(n), lens)
    valid[: len(win)] = (idx[: len(win)] + k) <= offsets[read_of[: len(win)] + 1]
    e = np.where(valid[: len(win)], np.exp(win), 0.0)
    ce = np.concatenate([[0.0], np.cumsum(e)])
    hi = np.minimum(offsets[1:], len(win))
    lo = np.minimum(offsets[:-1], len(win))
    return ce[hi] - ce[lo]


def vector_scores(qual, offsets, k, block=20000):
    """Expected number of error-free k-mers per read (the reference's sort key,
    get_sorted_fastq_for_cluster.py:23-33,150-152), vectorised with log-sums; only used to put the
    synthetic reads into the order the reference's sort stage would."""
    n = len(offsets) - 1
    out = np.zeros(n)
    for a in range(0, n, block):
        b = min(n, a + block)
        out[a:b] = _vector_scores_block(qual[offsets[a]:offsets[b]], offsets[a:b + 1] - offsets[a], k)
    return out


def make_workload(n_reads, seed, cache=True, n_species=10, with_templates=False, k=K, profile="ont", length=(700, 800),
                  per_read_len=None):
    """n_reads synthetic reads that pass the reference's quality filter, in score order.
    Returns (seq u8, qual u8, offsets i64, accessions list[str]) [+ templates list[str]]."""
    path = "/tmp/ngsid_bench_%d_%d_s%d_%s_k%d.npz" % (n_reads, seed, n_species, profile, k)
    if cache and os.path.exists(path):
        z = np.load(path, allow_pickle=False)
        out = (z["seq"], z["qual"], z["offsets"], [a.decode() for a in z["acc"]])
        return out + ([t.decode() for t in z["templates"]],) if with_templates else out
    from ngspeciesid_b200.synth import simulate_reads
    gen = int(n_reads * 1.12) + 64
    rs = simulate_reads(gen, n_species=n_species, len_lo=length[0], len_hi=length[1], seed=seed, profile=profile,
                        per_read_len=per_read_len)
    lens = rs.lengths()
    # quality filter of the sort stage (mean uncapped error probability, Q > 7)
    pu = 10.0 ** (-(rs.qual.astype(np.float64) - 33.0) / 10.0)
    cs = np.concatenate([[0.0], np.cumsum(pu)])
    mean_err = (cs[rs.offsets[1:]] - cs[rs.offsets[:-1]]) / lens
    ok = np.nonzero(-10.0 * np.log10(mean_err) > 7.0)[0][:n_reads]
    if len(ok) < n_reads:
        raise RuntimeError("synthetic pool too small")
    score = vector_scores(rs.qual, rs.offsets, k)[ok]
    order = ok[np.argsort(-score, kind="stable")]
    score_sorted = np.sort(-score, kind="stable") * -1.0
    new_off = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(lens[order], out=new_off[1:])
    src = np.repeat(rs.offsets[order] - new_off[:-1], lens[order]) + np.arange(new_off[-1])
    seq, qual = rs.seq[src], rs.qual[src]
    acc = ["read%d species=%d strand=%s_%r" % (i, rs.species[i], "+-"[rs.strand[i]], float(s))
           for i, s in zip(order, score_sorted)]
    templates = [t.tobytes().decode() for t in rs.templates]
    if cache:
        tmp = "%s.%d.tmp.npz" % (path, os.getpid())          # atomic: other ranks may be polling for it
        np.savez(tmp, seq=seq, qual=qual, offsets=new_off, acc=np.array([a.encode() for a in acc]),
                 templates=np.array([t.encode() for t in templates]))
        os.replace(tmp, path)
    return (seq, qual, new_off, acc, templates) if with_templates else (seq, qual, new_off, acc)


def slice_reads(seq, qual, offsets, lo, hi):
    a, b = offsets[lo], offsets[hi]
    return seq[a:b], qual[a:b], offsets[lo:hi + 1] - a


def read_array(seq, qual, offsets, acc, lo, hi):
    out = []
    for i in range(lo, hi):
        a, b = offsets[i], offsets[i + 1]
        out.append((i, 0, acc[i], seq[a:b].tobytes().decode(), qual[a:b].tobytes().decode(),
                    float(acc[i].split("_")[-1])))
    return out


def batch_bounds(lens, world):
    from ngspeciesid_b200.multi_gpu import batch_bounds as bb
    return bb(lens, world)


def merge_rounds(eng, params, acc_rank, batch_results, n_batches):
    """log2 rounds of pairwise consecutive batch merges on ONE engine that holds every read
    (modules/parallelize.py:137-217); used by the tests of the `--t N` decomposition.
    batch_results: {batch index (1-based): sorted representative read ids}.
    Returns ({rep: winner} merges, final reps)."""
    merges = {}
    cur = dict(batch_results)
    while len(cur) > 1:
        nxt = {}
        keys = sorted(cur)
        for j in range(0, len(keys), 2):
            lo = cur[keys[j]]
            nb = j // 2 + 1
            if j + 1 >= len(keys):
                nxt[nb] = lo
                continue
            hi = cur[keys[j + 1]]                        # processed in score order = id order
            assign, _via, _st = eng.cluster(K, W, params["max_gap"], np.array(hi, dtype=np.int32), acc_rank,
                                            init_reps=np.array(lo, dtype=np.int32))
            keep = list(lo)
            for rid, a in zip(hi, assign):
                if a >= 0:
                    merges[rid] = int(a)
                else:
                    keep.append(rid)
            nxt[nb] = sorted(keep)
        cur = nxt
    return merges, cur[min(cur)]


# ------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "sampler": "one nvidia-smi poller, GPU of rank 0, 200 ms"}


# ------------------------------------------------------------------------------------ GPU arm
def pick_config(args, world):
    name = args.config if args.config != "auto" else "c1"
    cfg = dict(CONFIGS[name])
    if args.reads:
        cfg["reads_per_gpu"] = args.reads
    for key in ("abundance_ratio", "max_seqs", "racon_iter"):
        v = getattr(args, key)
        if v is not None:
            cfg[key] = v
    cfg["key"] = name
    return cfg


def run_ours(args):
    import torch
    import torch.distributed as dist
    from ngspeciesid_b200 import engine as E
    from ngspeciesid_b200 import multi_gpu as M
    from ngspeciesid_b200.modules import p_minimizers_shared

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus must equal WORLD_SIZE under torchrun")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p_emp = p_minimizers_shared.p_emp_for(K, W)
    max_gap = E.max_gap_table(p_emp, 0.1)
    eng, mg, ce, pe = E.Engine(local), E.Engine(local), E.Engine(local), E.Engine(local)
    eng2 = E.Engine(local)                 # second batch engine: double-buffered input of the end-to-end loop
    if world > 1:
        # the library's own communicator (NCCL inside libngsid.so); the id travels over torch.distributed
        uid = [E.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.nccl_init(uid[0], rank, world)
        for e in (mg, ce, pe, eng2):
            e.nccl_share(eng)
    env = dict(torch=torch, dist=dist, E=E, M=M, rank=rank, world=world, local=local, p_emp=p_emp, max_gap=max_gap,
               engines=(eng, mg, ce, pe), eng2=eng2)
    cfg = pick_config(args, world)
    result = _measure(args, cfg, env)
    if (world >= 8 or args.force_north_star) and args.config == "auto" and not args.no_north_star:
        # the north-star run itself: BASELINE.json configs[3], cluster + consensus end to end on the same ranks
        import copy
        a3 = copy.copy(args)
        a3.config, a3.steps, a3.warmup = "c3", max(1, min(args.steps, 3)), max(1, min(args.warmup, 2))
        a3.no_roofline = a3.no_cpu = a3.no_modules = a3.no_concurrent = True
        a3.reads = 0
        r3 = _measure(a3, pick_config(a3, world), env)
        if rank == 0:
            result["north_star_configs3"] = r3
    if rank == 0:
        print(json.dumps(result))
    if world > 1:
        dist.barrier()
        for e in (eng2, pe, ce, mg, eng):
            e.close()
        dist.destroy_process_group()


def _measure(args, cfg, env):
    """One workload on the ranks of `env`: clustering (resident + end to end), consensus of the final
    clusters, and on rank 0 the optional single-GPU legs. Returns the result dict on rank 0."""
    torch, dist, E, M = env["torch"], env["dist"], env["E"], env["M"]
    rank, world, local, p_emp, max_gap = env["rank"], env["world"], env["local"], env["p_emp"], env["max_gap"]
    eng, mg, ce, pe = env["engines"]
    eng2 = env["eng2"]
    K, W = cfg.get("k", 13), cfg.get("w", 20)            # shadow the module defaults for this workload
    if (K, W) != (13, 20):
        from ngspeciesid_b200.modules import p_minimizers_shared
        p_emp = p_minimizers_shared.p_emp_for(K, W)
        max_gap = E.max_gap_table(p_emp, 0.1)
    wl = dict(n_species=cfg["species"], k=K, profile=cfg.get("profile", "ont"), length=cfg.get("len", (700, 800)),
              per_read_len=cfg.get("per_read_len"))
    n_total = cfg["reads_per_gpu"] * world
    seed = args.seed + world - 1
    if world > 1:
        # rank 0 generates (or finds) the pool and leaves it in the cache; the others load it
        if rank == 0:
            make_workload(n_total, seed, **wl)
        dist.barrier()
    seq, qual, offsets, acc, templates = make_workload(n_total, seed, with_templates=True, **wl)

    # --t N semantics: N consecutive batches of the score-sorted list by cumulative nucleotides
    bounds = batch_bounds(np.diff(offsets), world)
    lo, hi = bounds[rank], bounds[rank + 1]
    n_mine = hi - lo
    s_seq, s_qual, s_off = slice_reads(seq, qual, offsets, lo, hi)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    h_seq, h_qual, h_off = pin(s_seq), pin(s_qual), pin(s_off)
    my_acc = acc[lo:hi]
    my_scores = [float(a.split("_")[-1]) for a in my_acc]
    del seq, qual                                     # every rank keeps its own batch only

    pipe = M.Pipeline(eng, mg, ce, pe, rank=rank, world=world, k=K, w=W, alt=eng2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        eng.sync()

    state = {}

    def engines_launches():
        return sum(e.launch_count() for e in (eng, eng2, mg, ce, pe))

    # ---- warm-up + resident timing. The steps run through Pipeline.cluster_stream: the local pass of step s + 1
    # (second batch engine, second host thread, no collective) overlaps the survivor exchange and the merge rounds
    # of step s; both batch engines hold the batch.
    eng.upload(h_seq, h_qual, h_off)
    eng2.upload(h_seq, h_qual, h_off)
    dev = {"k1": 0.0, "k0": 0.0, "cluster": 0.0, "k4": 0.0, "map": 0.0}

    def on_step(_s, roots, L):
        state["roots"] = roots
        for k_ in dev:
            dev[k_] += L["device_ms"][k_]

    # Overlapped steps pay while the merge phase is short: measured 33.3 vs 34.6 ms per step at N = 2, 39.4 vs 41 at
    # N = 4, but 50.6 vs <= 49.7 at N = 8, where a rank spends ~8 ms per step inside the collectives of three merge
    # rounds and its spinning NCCL kernels slow the local pass that runs next to them.
    use_stream = world <= 4 and not args.no_stream

    def resident_loop(n):
        if use_stream:
            pipe.cluster_stream(n, max_gap, my_acc, my_scores, lo, n_total, upload=None, tile_reads=args.tile, on_step=on_step)
        else:
            for s_ in range(n):
                roots = pipe.cluster(max_gap, my_acc, my_scores, lo, n_total, tile_reads=args.tile)
                on_step(s_, roots, pipe.last_local)

    resident_loop(max(2, args.warmup))
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    for e in (eng, eng2, mg, ce, pe):
        e.reset_launch_count()
    pipe.phase = {}
    for k_ in dev:
        dev[k_] = 0.0
    t0 = time.perf_counter()
    resident_loop(args.steps)
    barrier()
    dt = time.perf_counter() - t0
    launches = engines_launches()
    phase_res = dict(pipe.phase)
    # ---- end to end (pinned host buffers -> H2D -> kernels -> D2H): the same step fed from the host, input
    # double-buffered through the public API (Pipeline.prefetch): the H2D copy and the packing of batch n+1
    # run on the alternate engine under the clustering pass of batch n. Every step's transfer is inside the
    # timed region; the first one of the region is not overlapped with anything.
    up = (h_seq, h_qual, h_off)

    def e2e_loop(n):
        pipe.prefetch(up)
        for s_ in range(n):
            state["roots"] = pipe.cluster(max_gap, my_acc, my_scores, lo, n_total, tile_reads=args.tile, prefetched=True,
                                          then_prefetch=up if s_ + 1 < n else None)

    e2e_loop(2)
    barrier()
    pipe.phase = {}
    t1 = time.perf_counter()
    e2e_loop(args.steps)
    barrier()
    dt_e2e = time.perf_counter() - t1
    clocks = sampler.stop() if sampler else None
    phase_e2e = dict(pipe.phase)

    def over_ranks(d, steps=None):
        """{phase: (max, min) over ranks} of per-step milliseconds."""
        keys = sorted(d)
        v = torch.tensor([d[k_] * 1000.0 / (steps or args.steps) for k_ in keys], dtype=torch.float64, device="cuda")
        if world > 1:
            mx, mn = v.clone(), v.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        else:
            mx = mn = v
        return {k_: {"max": float(a), "min": float(b)} for k_, a, b in zip(keys, mx.tolist(), mn.tolist())}

    phases = {"resident": over_ranks(phase_res), "e2e": over_ranks(phase_e2e),
              "device": over_ranks({k_: v / 1000.0 for k_, v in dev.items()})}
    if world > 1:
        t = torch.tensor([dt, dt_e2e, float(launches)], dtype=torch.float64, device="cuda")
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dt, dt_e2e = float(mx[0]), float(mx[1])
        launches = int(sm[2])

    result = None
    st = pipe.local_stats
    n_final = len(pipe.ms.final_reps())
    if rank == 0:
        value = n_total * args.steps / dt
        e2e_v = n_total * args.steps / dt_e2e
        result = {
            "metric": "reads/sec clustered + consensus bp/sec, 750bp ONT amplicons, 1/2/4/8 B200",
            "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1000.0 / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32/int32 (f64 for error rates)", "data": "synthetic",
            "config": workload_config(cfg, world),
            "timing": "wall clock between device syncs over all K steps, max over ranks (the host orchestrates the greedy pass); "
                      + ("the local pass of step s + 1 overlaps the exchange + merge rounds of step s (Pipeline.cluster_stream); "
                         "per-phase milliseconds per step as max/min over ranks in `phases` (they overlap: their sum exceeds ms_per_step)"
                         if use_stream else
                         "one step after the other (N >= 8: overlapped steps measured slower); per-phase milliseconds per step as "
                         "max/min over ranks in `phases`"),
            "steps_overlapped": bool(use_stream),
            "phases": phases,
            "e2e": {"value": e2e_v, "unit": "reads/s",
                    "h2d_bytes_per_step": int(offsets[-1]) * 2 + 8 * (n_total + world) + 8 * n_total,
                    "d2h_bytes_per_step": int(n_total * 5)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "cluster_stats_rank0": st, "final_clusters": n_final, "merge_rounds": pipe.merge_rounds,
        }
        if st["align_cells"] and dev["k4"] > 0:
            result["k4_gcups_rank0"] = st["align_cells"] * args.steps / (dev["k4"] / 1000.0) / 1e9

    # ---- independent batches in flight on one GPU (what `--t N` on one GPU does, modules/parallelize.py; the
    # headline above stays one pass at a time so that it is the same quantity at every N)
    if rank == 0 and world == 1 and not args.no_concurrent:
        result["concurrent_batches"] = concurrent_leg(E, M, local, K, W, max_gap, (h_seq, h_qual, h_off), my_acc, my_scores,
                                                      n_total, args, mg)

    # ---- consensus of the FINAL clusters (NGSpeciesID:124-158), clusters sharded over the ranks
    if not args.no_consensus:
        pipe.phase = {}
        for e in (eng, mg, ce, pe):
            e.reset_launch_count()
        centers, info = pipe.consensus(cfg["abundance_ratio"], cfg["max_seqs"], cfg["racon_iter"])     # warm-up
        barrier()
        pipe.phase = {}
        for e in (ce, pe):
            e.__dict__.pop("poa_acc", None)
        csteps = max(1, min(args.steps, 2))
        l0 = engines_launches()
        tc = time.perf_counter()
        for _ in range(csteps):
            centers, info = pipe.consensus(cfg["abundance_ratio"], cfg["max_seqs"], cfg["racon_iter"])
        barrier()
        dtc = (time.perf_counter() - tc) / csteps
        cl = engines_launches() - l0
        cph = dict(pipe.phase)
        if world > 1:
            t = torch.tensor([dtc], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dtc = float(t[0])
            t = torch.tensor([float(cl)], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            cl = int(t[0])
        cphase = over_ranks(cph, csteps)
        # read bases consumed: every read of the draft step once, every read of the polishing step per round
        lens_mean = float(offsets[-1]) / n_total
        bases = (info["reads_draft"] + info["reads_polish"] * cfg["racon_iter"]) * lens_mean
        if rank == 0:
            from oracle import consensus_oracle as co     # checker only: distance of the results to the templates
            rc = lambda s: co.revcomp(s)
            dists = []
            for _n, _cid, cons in centers[:12]:
                best = min(min(co.edit_distance(cons, t_), co.edit_distance(cons, rc(t_))) for t_ in templates)
                dists.append(best / float(len(cons)))
            result["consensus"] = {
                "metric": "consensus bp/s (read bases consumed by the draft POA + %d polishing rounds / wall time), FINAL clusters"
                          % cfg["racon_iter"],
                "value": bases / dtc, "unit": "bp/s", "seconds_per_step": dtc,
                "clusters_selected": info["clusters_selected"], "centres_after_rc_merge": info["centres_after_rc_merge"],
                "reads_draft": info["reads_draft"], "reads_polish_per_round": info["reads_polish"],
                "config": "clusters >= abundance_ratio %.3f x reads after the merge rounds, --max_seqs_for_consensus %d, "
                          "--racon_iter %d, --rc_identity_threshold 0.9; clusters sharded over %d rank(s), reads moved by "
                          "ngsid_exchange_reads" % (cfg["abundance_ratio"], cfg["max_seqs"], cfg["racon_iter"], world),
                "phases": cphase, "gpu_launches": cl,
                "consensus_lengths": [len(c[2]) for c in centers][:12],
                "edit_distance_to_nearest_template_per_base": [round(d, 5) for d in dists]}
            # K5 (partial-order alignment) of rank 0's share: a latency-bound kernel, reported as cells per second and
            # as time per layer step (one dependency chain of graph rows per job)
            k5 = {"calls": 0, "jobs": 0, "cells": 0, "layer_steps": 0, "device_ms": 0.0, "host_ms": 0.0, "call_ms": 0.0}
            for e in (ce, pe):
                for k_, v_ in e.__dict__.get("poa_acc", {}).items():
                    k5[k_] += v_
            if k5["call_ms"] > 0:
                result["consensus"]["k5_rank0"] = {
                    "kernel": "k5r_layer_kernel (one launch per layer step over all jobs) + host graph updates",
                    "bound": "latency: serial chain of graph rows per job, one CTA per job (DESIGN.md 4.3)",
                    "calls_per_step": k5["calls"] / csteps, "jobs_per_step": k5["jobs"] / csteps,
                    "dp_cells_per_step": k5["cells"] / csteps, "layer_steps_per_step": k5["layer_steps"] / csteps,
                    "seconds_in_poa_per_step": k5["call_ms"] / 1e3 / csteps,
                    "gcups": k5["cells"] / (k5["call_ms"] / 1e3) / 1e9,
                    "ms_per_layer_step": k5["call_ms"] / max(1, k5["layer_steps"]),
                    "device_share": k5["device_ms"] / k5["call_ms"], "host_graph_share": k5["host_ms"] / k5["call_ms"]}
        # CPU baseline of the consensus leg: the oracle (restated spoa / racon) on whole clusters, one per core
        if rank == 0 and not args.no_cpu and info["clusters_selected"]:
            result["consensus"]["cpu_baseline"] = consensus_cpu_baseline(pipe, ce, info, cfg, lens_mean)

    # ---- the same 100 k reads through the drop-in modules API (tuples in, dicts out)
    if rank == 0 and world == 1 and not args.no_modules:
        result["e2e_modules"] = modules_leg(s_seq, s_qual, s_off, my_acc, p_emp, pipe, local)

    # ---- K1 roofline on a replicated input far larger than L2 (rank 0 only)
    if rank == 0 and not args.no_roofline:
        rep = max(1, int(args.roofline_reads // max(1, n_mine)))
        big_seq = np.tile(s_seq, rep); big_qual = np.tile(s_qual, rep)
        blens = np.tile(np.diff(s_off), rep)
        big_off = np.zeros(len(blens) + 1, dtype=np.int64)
        np.cumsum(blens, out=big_off[1:])
        eng.upload(big_seq, big_qual, big_off)
        eng.minimizers_timed(K, W, 3)
        ms = eng.minimizers_timed(K, W, 10)
        _lc, counts, _k, _p = eng.get_minimizers(0, n_mine)
        nm_total = int(counts.sum()) * rep
        alg_bytes = int(((blens + 3) // 4).sum()) + 8 * len(blens) + 8 * nm_total + 4 * len(blens)
        peak, peak_src = 6650.0, "fallback"
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]); peak_src = "measured"
        except Exception:
            pass
        achieved = alg_bytes / (ms / 1000.0) / 1e9
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
            if int(tj["reads_per_launch"]) == len(blens) and (W - K + 1 == 8 and K <= 13):
                traffic = int(tj["dram_bytes_read"]) + int(tj["dram_bytes_write"]); traffic_src = tj["source"]
        except Exception:
            pass
        result["roofline"] = {"kernel": "k1_stream_kernel" if (W - K + 1 == 8 and K <= 13) else "k1_minimizers_kernel (generic, warp per read)", "bound": "hbm", "achieved": achieved, "peak": peak,
                              "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                              "traffic_source": traffic_src, "algorithmic_bytes_per_launch": int(alg_bytes),
                              "bound_note": "reported against HBM as BASELINE asks; the kernel is ALU-issue bound (DESIGN.md 4.0/4.1)",
                              "reads_per_launch": int(len(blens)), "bytes_per_read": alg_bytes / len(blens),
                              "kernel_ms": ms, "reads_per_s": len(blens) / (ms / 1000.0)}
        eng.upload(h_seq, h_qual, h_off)

    # ---- CPU baseline: the unmodified reference on a bounded prefix of the same ordered workload
    if rank == 0 and world == 1 and not args.no_cpu:
        from baseline import reference_arm
        cores = os.cpu_count() or 1
        cb = reference_arm.measure(s_seq, s_qual, s_off, my_acc, K, W, 1, 1, 0, cores, budget_s=args.cpu_budget, tag="cpu")
        result["cpu_baseline"] = {"value": cb["value"], "unit": "reads/s", "cores": cb["cores"], "kind": cb["kind"],
                                  "sample": cb["sample"], "one_core_reads_per_s": cb["one_copy_reads_per_s"],
                                  "aligner_seconds_share": cb["aligner_seconds_share"]}
        # parity on the sample: the oracle (pinned to the reference's golden vectors) on the same prefix
        from oracle import cluster_oracle as oc
        ns = min(cb["sample_reads"], n_mine, 4000)
        ra = read_array(s_seq, s_qual, s_off, my_acc, 0, ns)
        stats = oc.Stats()
        oc.single_clustering(ra, p_emp, oc.default_args(), stats)
        exp = [w_ for _r, w_, _h in stats.trace]
        result["parity_sample_identical"] = bool([int(x) for x in pipe.local_assign[:ns]] == exp)
    return result


def workload_config(cfg, world):
    shape = "750 bp ONT-error" if cfg.get("profile", "ont") == "ont" else "500-2000 bp PacBio-profile"
    return {"workload": "%s: %d synthetic %s reads per GPU (%d in total), %d species, k=%d w=%d, --t %d semantics"
                        % (cfg["name"], cfg["reads_per_gpu"], shape, cfg["reads_per_gpu"] * world, cfg["species"],
                           cfg.get("k", 13), cfg.get("w", 20), world),
            "reads_per_gpu": cfg["reads_per_gpu"], "total_reads": cfg["reads_per_gpu"] * world, "species": cfg["species"],
            "l2": "inputs_exceed_l2 (ASCII + packed reads + minimizer records = %.0f MB per GPU)"
                  % (cfg["reads_per_gpu"] * (750 * 2.25 + 119 * 8) / 1e6)}


def concurrent_leg(E, M, device, k, w, max_gap, up, accs, scores, n_total, args, mg):
    """Throughput with 2 and 3 independent passes in flight on the one GPU: every lane is a host thread with its own
    engine (context + stream) running the same step as the headline (K1 + K0 + greedy pass on its resident batch)."""
    import threading
    lanes_max = 3
    engs = [E.Engine(device) for _ in range(lanes_max)]
    pipes = [M.Pipeline(e, mg, None, None, rank=0, world=1, k=k, w=w) for e in engs]
    for e, p in zip(engs, pipes):
        e.upload(*up)
        p.cluster(max_gap, accs, scores, 0, n_total, tile_reads=args.tile)
    steps = max(2, args.steps // 2)
    out = {"metric": "reads/s with independent batches in flight on one GPU (one host thread, context and stream per batch)",
           "unit": "reads/s", "steps_per_lane": steps}

    def run(n_lanes):
        def work(p):
            for _ in range(steps):
                p.cluster(max_gap, accs, scores, 0, n_total, tile_reads=args.tile)
            p.eng.sync()
        th = [threading.Thread(target=work, args=(pipes[i],)) for i in range(n_lanes)]
        t = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        return time.perf_counter() - t
    for n_lanes in (1, 2, 3):
        run(n_lanes) if n_lanes > 1 else None                       # warm-up of the combination
        dt = run(n_lanes)
        out[str(n_lanes)] = n_total * steps * n_lanes / dt
    for e in engs:
        e.close()
    return out


def _consensus_cpu_worker(job):
    from oracle import consensus_oracle as co
    recs, iters = job
    t0 = time.perf_counter()
    d = co.spoa_consensus(recs)
    p = co.racon_polish(d, recs, iters, both_strands=False)
    return time.perf_counter() - t0, len(p)


def consensus_cpu_baseline(pipe, ce, info, cfg, lens_mean):
    """oracle/poa_oracle.cpp + consensus_oracle.py (the restated spoa / racon; -O3, scalar) on WHOLE clusters
    of the draft step, one cluster per host core at the same time."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    seq, qual, off = ce.h_seq, ce.h_qual, ce.offsets
    n_have = len(off) - 1
    per = cfg["max_seqs"] if cfg["max_seqs"] > 0 else 200
    jobs = []
    for j in range(min(cores, max(1, n_have // per))):
        ids = range(j * per, min(n_have, (j + 1) * per))
        jobs.append(([(seq[off[i]:off[i + 1]].tobytes().decode(), qual[off[i]:off[i + 1]].tobytes().decode()) for i in ids],
                     cfg["racon_iter"]))
    if not jobs:
        return None
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(len(jobs)) as pool:
        res = pool.map(_consensus_cpu_worker, jobs)
    dt = time.perf_counter() - t0
    bases = sum(sum(len(r[0]) for r in j[0]) for j in jobs) * (1 + cfg["racon_iter"])
    return {"value": bases / dt, "unit": "bp/s", "cores": len(jobs), "kind": "port",
            "sample": "%d whole read sets of %d reads (draft + %d polishing rounds each), one per core, oracle/poa_oracle.cpp "
                      "-O3 scalar + consensus_oracle.py; spoa / racon themselves are not in this image"
                      % (len(jobs), per, cfg["racon_iter"]),
            "seconds": dt, "slowest_job_seconds": max(r[0] for r in res)}


def modules_leg(s_seq, s_qual, s_off, my_acc, p_emp, pipe, device):
    """The call a user of the reference makes: modules.parallelize.single_clustering(read_array, p_emp_probs,
    args) -- list of tuples in, (clusters, representatives) dicts out -- on the same reads."""
    from ngspeciesid_b200.modules import parallelize
    from baseline import reference_arm
    n = len(my_acc)
    ra = read_array(s_seq, s_qual, s_off, my_acc, 0, n)
    a = reference_arm.reference_args(K, W, 1, None)      # the reference's CLI defaults
    a.device = device
    parallelize.single_clustering(list(ra), p_emp, a)              # warm-up
    steps = 3
    t0 = time.perf_counter()
    for _ in range(steps):
        clusters, reps = parallelize.single_clustering(list(ra), p_emp, a)
    dt = (time.perf_counter() - t0) / steps
    # same clustering as the array-level path
    exp = {}
    for i, r in enumerate(pipe.local_assign):
        exp.setdefault(int(r) if r >= 0 else i, []).append(my_acc[i])
    same = {k_: v for k_, v in clusters.items()} == exp
    return {"metric": "reads/s through modules.parallelize.single_clustering (Python tuples in, dicts out)",
            "value": n / dt, "unit": "reads/s", "seconds_per_call": dt, "reads": n,
            "clusters_equal_array_path": bool(same), "clusters": len(clusters)}


# ------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from baseline import reference_arm
    world = max(1, world, args.gpus if "WORLD_SIZE" not in os.environ else 1)
    cfg = pick_config(args, world)
    cores = os.cpu_count() or 1
    n_total = cfg["reads_per_gpu"] * world
    # a prefix of the score-ordered workload is a score-ordered workload: generate the prefix pool only
    n_pool = min(n_total, 60000)
    seq, qual, offsets, acc = make_workload(n_pool, args.seed + world - 1, n_species=cfg["species"])
    r = reference_arm.measure(seq, qual, offsets, acc, K, W, world, args.steps, args.warmup, cores,
                              budget_s=args.ref_budget, tag="arm")
    line = {
        "impl": "reference", "metric": "reads/sec clustered + consensus bp/sec, 750bp ONT amplicons, 1/2/4/8 B200",
        "value": r["value"], "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "python int/float", "data": "synthetic",
        "config": workload_config(cfg, world),
        "cpu_baseline": {"value": r["value"], "unit": "reads/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                         "t": r["t"], "copies": r["copies"], "sample_reads": r["sample_reads"],
                         "one_copy_reads_per_s": r["one_copy_reads_per_s"],
                         "aligner_seconds_share": r["aligner_seconds_share"], "clusters": r["clusters"],
                         "pool_note": "the sample is a prefix of a %d-read pool generated like the GPU arm's (same generator, "
                                      "seed and species; the GPU arm's pool has %d reads)" % (n_pool, n_total)},
        "e2e": {"value": r["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="auto", choices=["auto", "c1", "c3", "c4"])
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU (default: the configuration's)")
    ap.add_argument("--seed", type=int, default=1002)
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--cpu-budget", dest="cpu_budget", type=float, default=15.0)
    ap.add_argument("--ref-budget", dest="ref_budget", type=float, default=150.0)
    ap.add_argument("--roofline-reads", dest="roofline_reads", type=int, default=2000000)
    ap.add_argument("--no-roofline", dest="no_roofline", action="store_true")
    ap.add_argument("--no-cpu", dest="no_cpu", action="store_true")
    ap.add_argument("--no-consensus", dest="no_consensus", action="store_true")
    ap.add_argument("--no-modules", dest="no_modules", action="store_true")
    ap.add_argument("--no-concurrent", dest="no_concurrent", action="store_true", help="skip the batches-in-flight leg (N = 1)")
    ap.add_argument("--no-stream", dest="no_stream", action="store_true",
                    help="one step after the other instead of Pipeline.cluster_stream (the default from N = 8 on)")
    ap.add_argument("--no-north-star", dest="no_north_star", action="store_true", help="N = 8: skip the configs[3] run")
    ap.add_argument("--force-north-star", dest="force_north_star", action="store_true",
                    help="run the nested configs[3] workload (125 k reads per GPU) at any N: a dry run of the N = 8 path")
    ap.add_argument("--abundance-ratio", dest="abundance_ratio", type=float, default=None)
    ap.add_argument("--max-seqs", dest="max_seqs", type=int, default=None, help="--max_seqs_for_consensus of the consensus leg")
    ap.add_argument("--racon-iter", dest="racon_iter", type=int, default=None)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            import __graft_entry__ as g
            g.build()                                 # up to date on the GPU box (the built library travels)
        run_ours(args)


if __name__ == "__main__":
    main()
