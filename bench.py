#!/usr/bin/env python
"""
bench.py -- reads/s clustered on synthetic 750 bp ONT amplicon reads (BASELINE.json config 1:
100k reads, k=13, w=20, cluster-only) on N B200s.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the CPU path (oracle port) on the host cores

One "step" = one pass of the clustering hot path (K1 minimizers + K0 quality statistics + the
greedy pass with K2/K3 mapping and K4 block alignment) over the whole batch.
  value  : reads/s with the reads already resident in HBM (ASCII + packed), device-timed
  e2e    : the same through the host-buffer API (pinned host arrays -> H2D -> kernels -> D2H)
  roofline: K1 (minimizer extraction) timed alone with CUDA events on a replicated input that is
           far larger than L2, algorithmic bytes = packed read + (offset,len) + 8 B/minimizer + count
N > 1 follows the reference's --t N semantics (modules/parallelize.py): the score-sorted reads are
split into N consecutive batches, one per GPU, clustered independently, then log2(N) merge rounds
exchange representatives (ids only; every rank holds the synthetic pool). Weak scaling: the pool
has N x reads_per_gpu reads.
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


# ------------------------------------------------------------------------------------ workload
def _vector_scores_block(qual, offsets, k):
    p = np.minimum(10.0 ** (-(qual.astype(np.float64) - 33.0) / 10.0), 0.79433)
    lg = np.log1p(-p)
    cs = np.concatenate([[0.0], np.cumsum(lg)])
    n = len(offsets) - 1
    win = cs[k:] - cs[:-k]                     # window starting at flat position i
    lens = np.diff(offsets)
    valid = np.zeros(len(qual), dtype=bool)
    # windows that stay inside one read
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
    get_sorted_fastq_for_cluster.py:23-33,150-152), vectorised with log-sums; used only to put the
    synthetic reads in the order the reference's sort stage would. Blocks of reads keep the float64
    temporaries small (a 900 k-read pool would otherwise need ~40 GB per process)."""
    n = len(offsets) - 1
    out = np.zeros(n)
    for a in range(0, n, block):
        b = min(n, a + block)
        out[a:b] = _vector_scores_block(qual[offsets[a]:offsets[b]], offsets[a:b + 1] - offsets[a], k)
    return out


def make_workload(n_reads, seed, cache=True):
    """n_reads synthetic ONT reads that pass the reference's quality filter, in score order.
    Returns (seq u8, qual u8, offsets i64, accessions list[str])."""
    path = "/tmp/ngsid_bench_%d_%d.npz" % (n_reads, seed)
    if cache and os.path.exists(path):
        z = np.load(path, allow_pickle=False)
        return z["seq"], z["qual"], z["offsets"], [a.decode() for a in z["acc"]]
    from ngspeciesid_b200.synth import simulate_reads
    gen = int(n_reads * 1.12) + 64
    rs = simulate_reads(gen, n_species=10, len_lo=700, len_hi=800, seed=seed)
    lens = rs.lengths()
    # quality filter of the sort stage (mean uncapped error probability, Q > 7)
    pu = 10.0 ** (-(rs.qual.astype(np.float64) - 33.0) / 10.0)
    cs = np.concatenate([[0.0], np.cumsum(pu)])
    mean_err = (cs[rs.offsets[1:]] - cs[rs.offsets[:-1]]) / lens
    ok = np.nonzero(-10.0 * np.log10(mean_err) > 7.0)[0][:n_reads]
    if len(ok) < n_reads:
        raise RuntimeError("synthetic pool too small")
    score = vector_scores(rs.qual, rs.offsets, K)[ok]
    order = ok[np.argsort(-score, kind="stable")]
    score_sorted = np.sort(-score, kind="stable") * -1.0
    new_off = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(lens[order], out=new_off[1:])
    src = np.repeat(rs.offsets[order] - new_off[:-1], lens[order]) + np.arange(new_off[-1])
    seq, qual = rs.seq[src], rs.qual[src]
    acc = ["read%d species=%d strand=%s_%r" % (i, rs.species[i], "+-"[rs.strand[i]], float(s))
           for i, s in zip(order, score_sorted)]
    if cache:
        tmp = "%s.%d.tmp.npz" % (path, os.getpid())          # atomic: other ranks may be polling for it
        np.savez(tmp, seq=seq, qual=qual, offsets=new_off, acc=np.array([a.encode() for a in acc]))
        os.replace(tmp, path)
    return seq, qual, new_off, acc


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


# ------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ GPU arm
def merge_rounds(eng, params, acc_rank, batch_results, n_batches):
    """log2 rounds of pairwise consecutive batch merges (modules/parallelize.py:137-217), given
    every batch's surviving representatives. batch_results: {batch index (1-based): sorted list of
    representative read ids (global = uploaded index)}. Returns ({rep: winner} merges, final reps)."""
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


def batch_bounds(lens, world):
    """Read index bounds of the `world` consecutive batches of the score-sorted list
    (modules/parallelize.py:54-67: cut after the read that fills int(total_nt / N) + 1 nucleotides)."""
    n_total = len(lens)
    bounds = [0]
    if world > 1:
        limit = int(lens.sum() / world) + 1
        csum = np.cumsum(lens)
        base = 0
        while len(bounds) < world:
            j = int(np.searchsorted(csum, base + limit, side="left"))
            if j >= n_total:
                break
            bounds.append(j + 1)
            base = int(csum[j])
    while len(bounds) < world + 1:
        bounds.append(n_total)
    return bounds


def merge_representatives(eng2, seq, qual, offsets, acc, gathered, params):
    """gathered[b] = global read ids of the representatives batch b ended with. Uploads these reads
    only, runs the merge rounds and returns ({merged representative: winner}, final representatives),
    both in global read ids."""
    from ngspeciesid_b200 import engine as E
    ids = sorted(set(x for g in gathered for x in g))
    if not ids:
        return {}, []
    idx = {g: i for i, g in enumerate(ids)}
    parts = [slice_reads(seq, qual, offsets, g, g + 1) for g in ids]
    m_seq = np.concatenate([p[0] for p in parts]); m_qual = np.concatenate([p[1] for p in parts])
    m_off = np.zeros(len(ids) + 1, dtype=np.int64)
    np.cumsum([len(p[0]) for p in parts], out=m_off[1:])
    eng2.upload(m_seq, m_qual, m_off)
    eng2.minimizers(K, W)
    eng2.quality_stats()
    ar = E.accession_ranks([acc[g] for g in ids])
    merges, final = merge_rounds(eng2, params, ar, {b + 1: [idx[x] for x in g] for b, g in enumerate(gathered)}, len(gathered))
    return {ids[a]: ids[b] for a, b in merges.items()}, [ids[i] for i in final]


def run_ours(args):
    import torch
    import torch.distributed as dist
    from ngspeciesid_b200 import engine as E
    from ngspeciesid_b200.modules import p_minimizers_shared

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus must equal WORLD_SIZE under torchrun")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n_total = args.reads * world
    if world > 1:
        # rank 0 generates (or finds) the pool and leaves it in the cache; the others load it
        if rank == 0:
            make_workload(n_total, args.seed + world - 1)
        dist.barrier()
    seq, qual, offsets, acc = make_workload(n_total, args.seed + world - 1)
    p_emp = p_minimizers_shared.p_emp_for(K, W)
    params = {"max_gap": E.max_gap_table(p_emp, 0.1)}

    # --t N semantics: N consecutive batches of the score-sorted list by cumulative nucleotides
    bounds = batch_bounds(np.diff(offsets), world)
    lo, hi = bounds[rank], bounds[rank + 1]
    n_mine = hi - lo

    eng = E.Engine(local)
    s_seq, s_qual, s_off = slice_reads(seq, qual, offsets, lo, hi)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    h_seq, h_qual, h_off = pin(s_seq), pin(s_qual), pin(s_off)
    my_acc = acc[lo:hi]
    acc_rank = E.accession_ranks(my_acc)
    order = np.arange(n_mine, dtype=np.int32)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        eng.sync()

    state = {}

    def step(e2e):
        if e2e:
            eng.upload(h_seq, h_qual, h_off)
        eng.minimizers(K, W)
        eng.quality_stats()
        assign, via, st = eng.cluster(K, W, params["max_gap"], order, acc_rank, tile_reads=args.tile)
        state["assign"], state["via"], state["stats"] = assign, via, st
        if world > 1:
            # exchange surviving representatives (global ids) and run the merge rounds
            reps = [int(lo + i) for i in np.nonzero(assign == -1)[0]]
            gathered = [None] * world
            dist.all_gather_object(gathered, reps)
            state["gathered"] = gathered
        return assign

    def merge_on_rank0():
        """Merge rounds over the representatives of all batches (few hundred reads at most):
        run once on rank 0 inside the timed region of every step."""
        eng2 = state.setdefault("eng2", E.Engine(local))
        return merge_representatives(eng2, seq, qual, offsets, acc, state["gathered"], params)

    def full_step(e2e):
        step(e2e)
        if world > 1:
            if rank == 0:
                state["merge"] = merge_on_rank0()
            dist.barrier()

    # ---- warm-up + resident timing (device events on the engine's stream, max over ranks)
    eng.upload(h_seq, h_qual, h_off)
    for _ in range(args.warmup):
        full_step(False)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    eng.reset_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    phase = {"k1": 0.0, "k0": 0.0, "cluster": 0.0, "k4": 0.0, "map": 0.0}
    for _ in range(args.steps):
        full_step(False)
        phase["k1"] += eng.phase_ms(1); phase["k0"] += eng.phase_ms(2); phase["cluster"] += eng.phase_ms(3)
        phase["k4"] += eng.phase_ms(4); phase["map"] += eng.phase_ms(5)
    barrier()
    dt = time.perf_counter() - t0
    launches = eng.launch_count() + (state["eng2"].launch_count() if "eng2" in state else 0)
    clocks = sampler.stop()
    # device-event time of the steps (sum of the phases; host orchestration gaps are inside `cluster`)
    dev_ms = (phase["k1"] + phase["k0"] + phase["cluster"])
    # ---- end-to-end (host buffers -> H2D -> kernels -> D2H), wall clock bracketed by syncs
    for _ in range(1):
        full_step(True)
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        full_step(True)
    barrier()
    dt_e2e = time.perf_counter() - t1

    if world > 1:
        t = torch.tensor([dt, dt_e2e, dev_ms / 1000.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dt_e2e, dev_s = [float(x) for x in t.tolist()]
        dev_ms = dev_s * 1000.0

    result = None
    if rank == 0:
        st = state["stats"]
        value = n_total * args.steps / dt
        e2e_v = n_total * args.steps / dt_e2e
        result = {
            "metric": "reads/sec clustered (750 bp ONT amplicons, k=13 w=20, cluster-only)",
            "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1000.0 / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32/int32 (doubles for error rates)", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[1]: %d synthetic 750 bp ONT-error reads per GPU, 10 species, "
                                   "k=13 w=20, cluster-only, --t %d semantics" % (args.reads, world),
                       "reads_per_gpu": args.reads, "total_reads": n_total, "timing": "wall clock between stream syncs "
                       "(host orchestrates the greedy pass); device-event sum reported as device_ms_per_step",
                       "l2": "inputs_exceed_l2 (ASCII+packed reads + minimizers = %.0f MB per GPU)" %
                             ((offsets[hi] - offsets[lo]) * 2.25 / 1e6 + n_mine * 119 * 8 / 1e6),
                       "tile_reads": args.tile or 65536},
            "device_ms_per_step": dev_ms / args.steps,
            "phase_ms_per_step": {k_: v / args.steps for k_, v in phase.items()},
            "e2e": {"value": e2e_v, "unit": "reads/s",
                    "h2d_bytes_per_step": int(h_seq.nbytes + h_qual.nbytes + h_off.nbytes + acc_rank.nbytes + order.nbytes) * world,
                    "d2h_bytes_per_step": int(n_mine * 5) * world},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "cluster_stats": st,
        }
        if st["align_cells"] and phase["k4"] > 0:
            result["k4_gcups"] = st["align_cells"] * args.steps / (phase["k4"] / 1000.0) / 1e9


    # ---- consensus leg (BASELINE.json configs[2] shape): draft POA + racon-style polish of every
    # cluster above the abundance cut-off, capped with the reference's own --max_seqs_for_consensus.
    # N > 1: clusters shard -- every rank polishes the clusters of its own batch (as they stand
    # before the cross-batch merge rounds); no data-path collective, the aggregate is the sum of the
    # bases over the slowest rank's time.
    if not args.no_consensus:
        from ngspeciesid_b200.modules import consensus as C
        assign = state["assign"]
        rep = np.where(assign >= 0, assign, np.arange(n_mine))
        cutoff = int(args.abundance_ratio * n_mine)
        ids, counts = np.unique(rep, return_counts=True)
        big = ids[counts >= cutoff]
        order_idx = np.argsort(rep, kind="stable")
        starts = np.searchsorted(rep[order_idx], big)
        lists = [order_idx[st:st + c][: args.max_seqs].tolist() for st, c in zip(starts, counts[counts >= cutoff])]
        lens_ = np.diff(s_off)
        used_bases = int(sum(int(lens_[l].sum()) for l in lists))

        def consensus_step():
            drafts, _nodes = C.draft_consensus_batch(eng, lists)
            return C.polish_batch(eng, drafts, lists, args.racon_iter)

        l0 = eng.launch_count()
        consensus_step()
        barrier()
        tc = time.perf_counter()
        csteps = max(1, min(args.steps, 2))
        for _ in range(csteps):
            cons = consensus_step()
        barrier()
        dtc_ = (time.perf_counter() - tc) / csteps
        mine = {"bases": used_bases, "seconds": dtc_, "clusters": len(lists), "reads": int(sum(len(l) for l in lists)),
                "launches": int(eng.launch_count() - l0)}
        allc = [mine]
        if world > 1:
            allc = [None] * world
            dist.all_gather_object(allc, mine)
        if rank == 0:
            dtc_ = max(x["seconds"] for x in allc)
            result["consensus"] = {
                "metric": "consensus bp/s (read bases consumed by draft POA + %d polish rounds / wall time)" % args.racon_iter,
                "value": sum(x["bases"] for x in allc) * (1 + args.racon_iter) / dtc_, "unit": "bp/s", "seconds_per_step": dtc_,
                "clusters": sum(x["clusters"] for x in allc), "reads_used": sum(x["reads"] for x in allc),
                "config": "clusters >= abundance_ratio %.3f x reads, --max_seqs_for_consensus %d, --racon_iter %d; host buffers in, "
                          "consensus strings out%s" % (args.abundance_ratio, args.max_seqs, args.racon_iter,
                                                       "; clusters of each batch on its own GPU, max over ranks" if world > 1 else ""),
                "consensus_lengths": [len(c) for c in cons][:8], "gpu_launches": sum(x["launches"] for x in allc)}
        if rank == 0 and not args.no_cpu and lists:
            from oracle import consensus_oracle as co
            sample = lists[0][: args.cpu_consensus_reads]
            recs = [(s_seq[s_off[i]:s_off[i + 1]].tobytes().decode(), s_qual[s_off[i]:s_off[i + 1]].tobytes().decode()) for i in sample]
            t3 = time.perf_counter()
            d0 = co.spoa_consensus(recs)
            p0 = co.racon_polish(d0, recs, args.racon_iter)
            dt3 = time.perf_counter() - t3
            sb = sum(len(r[0]) for r in recs)
            g_d, _ = C.draft_consensus_batch(eng, [sample])
            g_p = C.polish_batch(eng, g_d, [sample], args.racon_iter)[0]
            result["consensus"]["cpu_baseline"] = {
                "value": sb * (1 + args.racon_iter) / dt3, "unit": "bp/s", "cores": 1, "kind": "port",
                "sample": "first %d reads of the largest cluster, oracle/poa_oracle.cpp + consensus_oracle.py" % len(sample)}
            result["consensus"]["parity_sample_edit_distance"] = int(co.edit_distance(g_p, p0))

    # ---- sort stage in front of the path (SURVEY.md 8 f rank 1): scores on the GPU, stable sort on the host
    if rank == 0 and not args.no_consensus:
        eng.sort_scores(K)
        eng.sync()
        ts = time.perf_counter()
        for _ in range(3):
            sc, er = eng.sort_scores(K)
            srt_order = np.argsort(-sc, kind="stable")
        dts = (time.perf_counter() - ts) / 3
        result["sort_stage"] = {"metric": "reads/s scored + ordered (get_sorted_fastq_for_cluster arithmetic; reads resident, "
                                          "scores D2H, stable argsort on the host)", "value": n_mine / dts, "unit": "reads/s",
                                "ms": dts * 1e3, "reads": int(n_mine)}
        if not args.no_cpu:
            from oracle import cluster_oracle as oc
            ns_ = min(n_mine, 2000)
            t4 = time.perf_counter()
            ref_sc = [oc.expected_error_free_kmers_score(s_qual[s_off[i]:s_off[i + 1]].tobytes().decode(), K) for i in range(ns_)]
            dt4 = time.perf_counter() - t4
            result["sort_stage"]["cpu_baseline"] = {"value": ns_ / dt4, "unit": "reads/s", "cores": 1, "kind": "port",
                                                    "sample": "first %d reads, oracle/cluster_oracle.py" % ns_}
            result["sort_stage"]["parity_sample_identical"] = bool(ref_sc == [float(x) for x in sc[:ns_]])

    # ---- ingest in front of the sort stage (SURVEY.md 8 f rank 3): host C parser vs the generator
    if rank == 0 and not args.no_consensus:
        import io
        from ngspeciesid_b200.modules import help_functions as hf
        ni = min(n_mine, 20000)
        text = "".join("@%s\n%s\n+\n%s\n" % (my_acc[i], s_seq[s_off[i]:s_off[i + 1]].tobytes().decode(),
                                               s_qual[s_off[i]:s_off[i + 1]].tobytes().decode()) for i in range(ni))
        data = text.encode()
        t5 = time.perf_counter()
        fa = hf.parse_fastq_bytes(data)
        dt5 = time.perf_counter() - t5
        t6 = time.perf_counter()
        n_py = sum(1 for _ in hf.readfq(io.StringIO(text)))
        dt6 = time.perf_counter() - t6
        result["ingest"] = {"metric": "FASTQ bytes/s parsed into upload-ready arrays (ngsid_fastq_parse, host, 1 thread)",
                            "value": len(data) / dt5, "unit": "B/s", "reads": int(len(fa)), "bytes": len(data),
                            "cpu_baseline": {"value": len(data) / dt6, "unit": "B/s", "cores": 1, "kind": "port",
                                             "sample": "the same %d records through the readfq generator mirror" % n_py},
                            "parity_identical": bool(len(fa) == n_py == ni and (fa.seq == s_seq[:s_off[ni]]).all()
                                                     and (fa.qual == s_qual[:s_off[ni]]).all())}

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
        # DRAM traffic of one launch from the committed ncu --set full capture (same command, same size)
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
            if int(tj["reads_per_launch"]) == len(blens):
                traffic = int(tj["dram_bytes_read"]) + int(tj["dram_bytes_write"]); traffic_src = tj["source"]
        except Exception:
            pass
        result["roofline"] = {"kernel": "k1_stream_kernel (thread per read: compress + window minima; warp per 32 reads: output)", "bound": "hbm", "achieved": achieved, "peak": peak,
                              "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                              "algorithmic_bytes_per_launch": int(alg_bytes), "bound_note": "reported against HBM as BASELINE asks; the kernel is ALU-issue bound (DESIGN.md 4.0/4.1)",
                              "reads_per_launch": int(len(blens)), "bytes_per_read": alg_bytes / len(blens),
                              "kernel_ms": ms, "reads_per_s": len(blens) / (ms / 1000.0)}
        eng.upload(h_seq, h_qual, h_off)

    # ---- CPU baseline: the oracle port on a bounded prefix of the same ordered workload
    if rank == 0 and not args.no_cpu:
        from oracle import cluster_oracle as oc
        ns = min(args.cpu_sample, n_mine)
        ra = read_array(seq, qual, offsets, acc, 0, ns)
        stats = oc.Stats()
        t2 = time.perf_counter()
        oc.single_clustering(ra, p_emp, oc.default_args(), stats)
        dtc = time.perf_counter() - t2
        exp = [w_ for _r, w_, _h in stats.trace]
        got = [int(x) for x in state["assign"][:ns]] if world == 1 else None
        result["cpu_baseline"] = {"value": ns / dtc, "unit": "reads/s", "cores": 1, "kind": "port",
                                  "sample": "first %d reads of the same score-ordered workload, oracle/cluster_oracle.py "
                                            "(Python restatement + C aligner), single thread" % ns}
        if got is not None:
            result["parity_sample_identical"] = bool(got == exp)
    if rank == 0:
        print(json.dumps(result))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------ reference arm
def _ref_worker(job):
    from oracle import cluster_oracle as oc
    ra, p_emp, bi = job
    clusters = {r[0]: [r[2]] for r in ra}
    reps = {r[0]: tuple(r) for r in ra}
    res = oc.reads_to_clusters(clusters, reps, ra, p_emp, {}, bi, oc.default_args())
    return res[bi]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import cluster_oracle as oc
    from ngspeciesid_b200.modules import p_minimizers_shared
    cores = os.cpu_count() or 1
    n_total = args.reads * max(1, world)
    seq, qual, offsets, acc = make_workload(n_total, args.seed + max(1, world) - 1)
    p_emp = p_minimizers_shared.p_emp_for(K, W)
    # bounded sample: about two minutes of CPU work for the whole run whatever --steps is
    # (the port clusters ~300 reads/s per core)
    per_core = min(args.ref_reads_per_core, max(100, int(120 * 300 / max(1, args.steps))))
    ns = min(n_total, per_core * cores)
    ra = read_array(seq, qual, offsets, acc, 0, ns)
    a = oc.default_args(nr_cores=cores)
    batches = [b for b in oc.split_batches(ra, cores, "total_nt") if b]

    def one_step():
        with mp.get_context("fork").Pool(len(batches)) as pool:
            res = pool.map(_ref_worker, [(b, p_emp, i + 1) for i, b in enumerate(batches)])
        # merge rounds (tiny) in-process, as the reference does after joining the pool
        all_cl, all_rp, all_db = {}, {}, {}
        for c, r, d, bi in res:
            all_cl.update(c); all_rp.update(r); all_db[bi] = d
        arr = [(v[0], v[1], v[2], v[3], v[4], v[5]) for _, v in sorted(all_rp.items(), key=lambda x: x[1][5], reverse=True)]
        while True:
            groups = oc.pair_batches(arr)
            if len(groups) <= 1 and len(all_db) <= 1:
                break
            n_all_cl, n_all_rp, n_all_db = {}, {}, {}
            for gi, g in enumerate(groups):
                low = min(r[1] for r in g)
                cl = {r[0]: all_cl[r[0]] for r in g}; rp = {r[0]: all_rp[r[0]] for r in g}
                out = oc.reads_to_clusters(cl, rp, g, p_emp, all_db[low], gi + 1, a)[gi + 1]
                n_all_cl.update(out[0]); n_all_rp.update(out[1]); n_all_db[gi + 1] = out[2]
            all_cl, all_rp, all_db = n_all_cl, n_all_rp, n_all_db
            arr = [(v[0], v[1], v[2], v[3], v[4], v[5]) for _, v in sorted(all_rp.items(), key=lambda x: x[1][5], reverse=True)]
            if len(groups) == 1:
                break
        return len(all_cl)

    for _ in range(min(args.warmup, 1)):
        one_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step()
    dt = time.perf_counter() - t0
    v = ns * args.steps / dt
    sample = ("prefix of %d reads (%d per core) of the same score-ordered workload; oracle port of the reference's "
              "--t %d path: %d batches in a process pool + merge rounds" % (ns, per_core, cores, len(batches)))
    print(json.dumps({
        "impl": "reference", "metric": "reads/sec clustered (750 bp ONT amplicons, k=13 w=20, cluster-only)",
        "value": v, "unit": "reads/s", "n_gpus": max(1, world), "steps": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": dt * 1000.0 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "python int/float", "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[1]: %d synthetic 750 bp ONT-error reads per GPU, 10 species, k=13 w=20, "
                               "cluster-only" % args.reads, "sample_reads": ns},
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=100000, help="reads per GPU")
    ap.add_argument("--seed", type=int, default=1002)
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--cpu-sample", dest="cpu_sample", type=int, default=4000)
    ap.add_argument("--ref-reads-per-core", dest="ref_reads_per_core", type=int, default=1500)
    ap.add_argument("--roofline-reads", dest="roofline_reads", type=int, default=2000000)
    ap.add_argument("--no-roofline", dest="no_roofline", action="store_true")
    ap.add_argument("--no-cpu", dest="no_cpu", action="store_true")
    ap.add_argument("--no-consensus", dest="no_consensus", action="store_true")
    ap.add_argument("--abundance-ratio", dest="abundance_ratio", type=float, default=0.02)
    ap.add_argument("--max-seqs", dest="max_seqs", type=int, default=200, help="--max_seqs_for_consensus of the consensus leg")
    ap.add_argument("--racon-iter", dest="racon_iter", type=int, default=3)
    ap.add_argument("--cpu-consensus-reads", dest="cpu_consensus_reads", type=int, default=60)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        import __graft_entry__ as g
        g.build()
        run_ours(args)


if __name__ == "__main__":
    main()
