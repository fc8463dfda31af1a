"""
N-GPU driver of the two hot paths: the reference's `--t N` clustering (modules/parallelize.py:107-217)
with batch i on GPU i, and the consensus step (NGSpeciesID:124-158, modules/consensus.py:249-278,
148-183, 186-246) on the FINAL clusters, sharded by cluster over the GPUs. One process per GPU; all
bulk data (representatives with their minimizer records, reads of a cluster) moves device to device
through the library's NCCL data plane (csrc/nccl_plane.cuh); the host only moves plans: ids, sizes,
accession strings, consensus strings.

Semantics: the result of the clustering equals the reference run with `--t <world>` (batches by
cumulative nucleotides, log2 rounds of pairwise batch merges, modules/parallelize.py:33-81,137-217);
the pairs of a round are independent and run on different ranks. With one rank everything below
degenerates to `--t 1` without any collective.

The functions that only compute plans are pure (tested on the CPU over gloo with an oracle-backed
stand-in engine, tests/test_multi_gpu_gloo.py); `Pipeline` drives Engine objects.
"""
import struct
import threading
import time

import numpy as np

from . import engine as E

INT32_MIN = np.iinfo(np.int32).min


# ------------------------------------------------------------------------------------- pure planning
def batch_bounds(lens, n_batches):
    """Read-index bounds of the consecutive batches of the score-sorted list
    (modules/parallelize.py:54-67: a batch ends with the read that brings its nucleotides to
    int(total / n) + 1)."""
    n_total = len(lens)
    bounds = [0]
    if n_batches > 1:
        limit = int(int(lens.sum()) / n_batches) + 1
        csum = np.cumsum(lens)
        base = 0
        while len(bounds) < n_batches:
            j = int(np.searchsorted(csum, base + limit, side="left"))
            if j >= n_total:
                break
            bounds.append(j + 1)
            base = int(csum[j])
    while len(bounds) < n_batches + 1:
        bounds.append(n_total)
    return bounds


def pack_rep_tags(gids, scores, sizes, accs):
    out = [struct.pack("<q", len(gids))]
    for g, s, z, a in zip(gids, scores, sizes, accs):
        b = a.encode("utf-8")
        out.append(struct.pack("<qdqI", int(g), float(s), int(z), len(b)))
        out.append(b)
    return b"".join(out)


def unpack_rep_tags(blob):
    (n,) = struct.unpack_from("<q", blob, 0)
    o = 8
    gids, scores, sizes, accs = [], [], [], []
    for _ in range(n):
        g, s, z, ln = struct.unpack_from("<qdqI", blob, o)
        o += 28
        gids.append(g); scores.append(s); sizes.append(z); accs.append(blob[o:o + ln].decode("utf-8"))
        o += ln
    return gids, scores, sizes, accs


class MergeState(object):
    """Bookkeeping of the merge rounds over the gathered representatives g = 0..R-1 (rank order, then
    processing order = global score order). groups: {batch index: [g, ...]} as
    modules/parallelize.py:196-215 regroups the survivors; glist[g]: the round-0 clusters that the
    cluster of g consists of, in the reference's concatenation order (modules/cluster.py:338-345)."""

    def __init__(self, counts):
        self.R = int(sum(counts))
        self.groups, o = {}, 0
        for b, c in enumerate(counts):
            self.groups[b + 1] = list(range(o, o + int(c)))
            o += int(c)
        self.glist = {g: [g] for g in range(self.R)}
        self.merged_into = {}

    def pairs(self):
        """[(new batch index, lower group, upper group or None)] of the next round."""
        keys = sorted(self.groups)
        return [(j // 2 + 1, self.groups[keys[j]], self.groups[keys[j + 1]] if j + 1 < len(keys) else None)
                for j in range(0, len(keys), 2)]

    def apply(self, pairs, dec):
        """dec[g] for every g of an upper group: winner g or -1 (stays a representative)."""
        nxt = {}
        for nb, lo, hi in pairs:
            keep = list(lo)
            for h in hi or []:
                w = int(dec[h])
                if w >= 0:
                    self.merged_into[h] = w
                    self.glist[w].extend(self.glist.pop(h))
                else:
                    keep.append(h)
            nxt[nb] = sorted(keep)
        self.groups = nxt

    def done(self):
        return len(self.groups) <= 1

    def final_reps(self):
        return self.groups[min(self.groups)] if self.groups else []


def select_clusters(glist, size0, scores, n_total, abundance_ratio, max_seqs):
    """Clusters that get a consensus, in the reference's order (modules/consensus.py:254: size, then
    representative score, descending; ties keep the order of the final clusters dict), each with the
    round-0 clusters it takes reads from: [(root g, n_reads, [(g, take), ...])]."""
    cutoff = int(abundance_ratio * n_total)
    sizes = {r: sum(size0[g] for g in gl) for r, gl in glist.items()}
    order = sorted(sorted(glist), key=lambda r: (sizes[r], scores[r]), reverse=True)
    out = []
    for r in order:
        if sizes[r] < cutoff:
            continue
        left = sizes[r] if max_seqs < 0 else min(sizes[r], max_seqs)
        segs = []
        for g in glist[r]:
            if left <= 0:
                break
            t = min(size0[g], left)
            segs.append((g, t))
            left -= t
        out.append((r, sizes[r], segs))
    return out, sizes


def assign_owners(weights, world):
    """Largest first onto the least loaded rank (ties: lowest rank). Deterministic on every rank."""
    load = [0] * world
    owner = [0] * len(weights)
    for i in sorted(range(len(weights)), key=lambda i: (-weights[i], i)):
        r = min(range(world), key=lambda q: (load[q], q))
        owner[i] = r
        load[r] += weights[i]
    return owner


def merge_reverse_complements(n_reads, identity, threshold):
    """Which centres survive modules/consensus.py:148-183 and what they absorb, given
    identity[i][j] = highest_aln_identity(centre i, centre j) for j > i. Returns
    [(i, merged read count, [i, absorbed j, ...])]. The reference's inner loop also looks at centres
    that an earlier centre has already absorbed; so does this."""
    n = len(n_reads)
    gone = set()
    out = []
    for i in range(n):
        if i in gone:
            continue
        members, total = [i], n_reads[i]
        for j in range(i + 1, n):
            if identity[i][j] >= threshold:
                total += n_reads[j]
                gone.add(j)
                members.append(j)
        out.append((i, total, members))
    return out


# ------------------------------------------------------------------------------------- driver
class Pipeline(object):
    """eng: Engine of this rank's batch; mg / ce / pe: further Engines on the same GPU for the gathered
    representatives, the reads of the draft step and the reads of the polishing step."""

    def __init__(self, eng, mg, ce, pe, rank=0, world=1, k=13, w=20, cluster_kw=None, alt=None):
        self.eng, self.mg, self.ce, self.pe = eng, mg, ce, pe
        self.alt = alt                      # second batch engine on the same GPU: double-buffered input (prefetch)
        self.rank, self.world, self.k, self.w = rank, world, k, w
        self.kw = dict(cluster_kw or {})
        self.phase = {}
        self._pf, self._pf_time, self._pf_err = None, 0.0, None
        self._pf_delay = 0.0                 # seconds into the current pass at which a prefetch starts its transfer

    def _tick(self, name, t0):
        self.eng.sync()
        self.phase[name] = self.phase.get(name, 0.0) + (time.perf_counter() - t0)
        return time.perf_counter()

    # ---- double-buffered input -------------------------------------------------------------------
    def prefetch(self, upload, delay=0.0):
        """Start moving the NEXT batch in: H2D copy + 2-bit packing of upload = (seq, qual, offsets) on the
        alternate engine (its own stream), driven by a host thread, so that the transfer runs under the
        clustering pass of the current batch. `cluster(prefetched=True)` then switches to that engine and runs
        K1 / K0 there (kernels of a second stream would only get SM slots when the pass's long alignment launch
        ends, i.e. after the pass: measured 15 ms for 0.7 ms of work). delay: seconds to wait before the
        transfer starts (cluster(then_prefetch=...) passes the point where its pass enters the bulk phase)."""
        if self.alt is None or self._pf is not None:
            raise RuntimeError("prefetch needs an alternate engine and no prefetch in flight")
        alt = self.alt

        def work():
            try:
                # A large H2D transfer delays every small copy and sync of a pass that runs under it (measured:
                # +3 ms per 150 MB, scripts/prefetch_probe.py, with any copy code). The first third of a pass is its
                # latency-bound part (small speculation tiles); the rest waits for one long alignment launch. So the
                # transfer starts once the pass is in that bulk phase.
                if delay > 0:
                    time.sleep(delay)
                t = time.perf_counter()
                alt.upload(*upload)
                alt.sync()
                self._pf_time = time.perf_counter() - t
            except Exception as exc:            # surfaces in cluster()
                self._pf_err = exc
        self._pf = threading.Thread(target=work, daemon=True)
        self._pf.start()

    def _take_prefetched(self):
        if self._pf is None:
            raise RuntimeError("no prefetch in flight")
        self._pf.join()
        self._pf = None
        if self._pf_err is not None:
            err, self._pf_err = self._pf_err, None
            raise err
        self.eng, self.alt = self.alt, self.eng
        self.phase["upload_prefetched"] = self.phase.get("upload_prefetched", 0.0) + self._pf_time

    # ---- clustering ------------------------------------------------------------------------------
    def cluster(self, max_gap, accs, scores, gid0, n_total, upload=None, tile_reads=0, prefetched=False,
                then_prefetch=None):
        """accs / scores: accession strings (with score suffix) and scores of the local reads in
        processing order; gid0: global index of the first local read. upload = (seq, qual, offsets)
        host arrays, or None when the reads are already resident. prefetched=True: the batch was handed
        to `prefetch` before (alternate engine, already on the device); then_prefetch = the batch after this one,
        started right away so that its transfer runs under this pass.
        Returns the final root (global read id) of every local read (-2: skipped by the reference)."""
        t = time.perf_counter()
        if prefetched:
            self._take_prefetched()
            t = self._tick("wait_prefetch", t)
            if then_prefetch is not None:
                self.prefetch(then_prefetch, delay=self._pf_delay)
            t = self._tick("start_prefetch", t)
        L = self.local_pass(self.eng, max_gap, accs, scores, gid0, upload=None if prefetched else upload,
                            tile_reads=tile_reads, phase=self.phase)
        self._pf_delay = 0.35 * L["t_pass"]
        self.last_local = L                              # record of the local half (device timers, stats) of this step
        return self.exchange_and_merge(L, max_gap, n_total)

    def local_pass(self, eng, max_gap, accs, scores, gid0, upload=None, tile_reads=0, phase=None):
        """This rank's batch on `eng`, no communication: (upload,) K1 minimizers, K0 quality statistics, the
        greedy pass; then the plan of its survivors. Returns the record `exchange_and_merge` takes. May run on
        a host thread of its own (cluster_stream): it touches `eng` and the given phase dict only."""
        phase = self.phase if phase is None else phase

        def tick(name, t0):
            eng.sync()
            phase[name] = phase.get(name, 0.0) + (time.perf_counter() - t0)
            return time.perf_counter()
        t = time.perf_counter()
        if upload is not None:
            eng.upload(*upload)
            t = tick("upload", t)
        eng.minimizers(self.k, self.w)
        eng.quality_stats()
        t = tick("k1_k0", t)
        n = len(accs)
        if getattr(self, "_acc_rank_for", None) is not accs:
            self._acc_rank, self._acc_rank_for = E.accession_ranks(accs), accs
        assign, via, st = eng.cluster(self.k, self.w, max_gap, np.arange(n, dtype=np.int32), self._acc_rank,
                                      tile_reads=tile_reads, **self.kw)
        t_pass = time.perf_counter() - t
        t = tick("cluster_local", t)
        reps = np.nonzero(assign == -1)[0].astype(np.int32)
        rep_of = np.where(assign >= 0, assign, np.arange(n))
        rep_of[assign == -2] = -1
        size0_local = np.bincount(rep_of[rep_of >= 0], minlength=n)[reps]
        blob = pack_rep_tags([gid0 + int(r) for r in reps], [scores[r] for r in reps], size0_local, [accs[r] for r in reps])
        dev = {}
        if hasattr(eng, "phase_ms"):                     # device timers of this pass (stand-in engines of the CPU tests have none)
            dev = {"k1": eng.phase_ms(1), "k0": eng.phase_ms(2), "cluster": eng.phase_ms(3), "k4": eng.phase_ms(4), "map": eng.phase_ms(5)}
        return {"eng": eng, "n": n, "assign": assign, "stats": st, "reps": reps, "rep_of": rep_of, "blob": blob,
                "t_pass": t_pass, "device_ms": dev}

    def exchange_and_merge(self, L, max_gap, n_total):
        """The communication half of a step for the local result L: plans over the host all-gather, the survivors
        device to device into `mg`, the merge rounds (pairs of a round on different ranks), the final root of every
        local read. All collectives of a step are issued here, by the calling thread, in the same order on every rank."""
        eng, mg = L["eng"], self.mg
        n, reps, rep_of = L["n"], L["reps"], L["rep_of"]
        self.eng = eng                                   # the engine that holds this step's batch (consensus reads from it)
        self.local_assign, self.local_stats = L["assign"], L["stats"]
        t = time.perf_counter()
        if self.world > 1:
            blobs = eng.allgather_bytes(L["blob"])
            counts = eng.gather_representatives(reps, mg)
        else:
            # one rank: there is no merge round, nothing ever reads a gathered copy of the representatives
            blobs, counts = [L["blob"]], np.array([len(reps)], dtype=np.int64)
        t = self._tick("gather_representatives", t)
        gids, gscores, gsizes, gaccs = [], [], [], []
        for b_ in blobs:
            a_, s_, z_, c_ = unpack_rep_tags(b_)
            gids += a_; gscores += s_; gsizes += z_; gaccs += c_
        assert [len(unpack_rep_tags(b_)[0]) for b_ in blobs] == [int(c_) for c_ in counts]
        self.g_gid, self.g_score, self.g_size0, self.g_acc = gids, gscores, gsizes, gaccs
        self.g_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        ms = MergeState(counts)
        g_rank = E.accession_ranks(gaccs)
        rounds = 0
        while not ms.done():
            pairs = ms.pairs()
            dec = np.full(ms.R, INT32_MIN, dtype=np.int32)
            todo = [p_ for p_ in pairs if p_[2]]
            for pj, (_nb, lo, hi) in enumerate(todo):
                if pj % self.world != self.rank:
                    continue
                a_, _v, _s = mg.cluster(self.k, self.w, max_gap, np.asarray(hi, dtype=np.int32), g_rank,
                                        init_reps=np.asarray(lo, dtype=np.int32), **self.kw)
                dec[np.asarray(hi)] = np.where(a_ >= 0, a_, -1)
            eng.allreduce(dec, "max")
            ms.apply(pairs, dec)
            rounds += 1
        self.merge_rounds = rounds
        self.ms = ms
        t = self._tick("merge_rounds", t)
        # ---- final root of every local read
        root_of_g = np.arange(ms.R)
        for g in range(ms.R):
            r = g
            while r in ms.merged_into:
                r = ms.merged_into[r]
            root_of_g[g] = r
        self.root_of_g = root_of_g
        g_of_local = np.full(n, -1, dtype=np.int64)
        g_of_local[reps] = self.g_off[self.rank] + np.arange(len(reps))
        self.local_rep_of = rep_of
        self.local_reps = reps
        out = np.full(n, -2, dtype=np.int64)
        ok = rep_of >= 0
        out[ok] = np.asarray(gids, dtype=np.int64)[root_of_g[g_of_local[rep_of[ok]]]]
        self.n_total = n_total
        return out

    def cluster_stream(self, n_steps, max_gap, accs, scores, gid0, n_total, upload=None, tile_reads=0, on_step=None):
        """n_steps batches one after the other with the two halves of a step overlapped: the local pass of step
        s + 1 runs on the alternate batch engine, driven by a second host thread, while this thread does the
        exchange and the merge rounds of step s (and waits there for slower ranks). The second thread issues no
        collective. upload = (seq, qual, offsets) fed to every step (end to end), or None when BOTH batch engines
        hold the batch already. on_step(step, roots, local_record) is called after every step.
        Returns the roots of the last step; `self.eng` is then the engine that holds its batch."""
        import queue
        if self.alt is None:
            raise RuntimeError("cluster_stream needs an alternate engine")
        engs = [self.eng, self.alt]
        free = [threading.Semaphore(1), threading.Semaphore(1)]
        q = queue.Queue()
        wphase = {}

        def worker():
            try:
                for s_ in range(n_steps):
                    free[s_ & 1].acquire()
                    q.put(self.local_pass(engs[s_ & 1], max_gap, accs, scores, gid0, upload=upload, tile_reads=tile_reads, phase=wphase))
            except BaseException as exc:                 # surfaces in the calling thread
                q.put(exc)
        th = threading.Thread(target=worker, daemon=True)
        th.start()
        roots = None
        for s_ in range(n_steps):
            t = time.perf_counter()
            L = q.get()
            if isinstance(L, BaseException):
                raise L
            self.phase["wait_local_pass"] = self.phase.get("wait_local_pass", 0.0) + (time.perf_counter() - t)
            roots = self.exchange_and_merge(L, max_gap, n_total)
            free[s_ & 1].release()
            if on_step is not None:
                on_step(s_, roots, L)
        th.join()
        for k_, v_ in wphase.items():
            self.phase[k_] = self.phase.get(k_, 0.0) + v_
        self.alt = engs[0] if self.eng is engs[1] else engs[1]
        return roots

    # ---- consensus on the final clusters -----------------------------------------------------------
    def consensus(self, abundance_ratio, max_seqs, racon_iter, rc_identity_threshold=0.9):
        """Draft (spoa-equivalent), reverse-complement merge and `racon_iter` polishing rounds of every
        final cluster above the abundance cut-off. Returns on every rank
        [[n_reads, c_id (global read id of the representative), polished consensus], ...] in the
        reference's order, and a dict of counters."""
        from .modules import consensus as C
        eng, ce, pe = self.eng, self.ce, self.pe
        t = time.perf_counter()
        sel, _sizes = select_clusters(self.ms.glist, self.g_size0, self.g_score, self.n_total, abundance_ratio, max_seqs)
        used = [sum(tk for _g, tk in segs) for _r, _n, segs in sel]
        owner = assign_owners(used, self.world)
        # local reads of every local round-0 cluster, representative first
        order = np.argsort(self.local_rep_of, kind="stable")
        order = order[self.local_rep_of[order] >= 0]
        starts = np.searchsorted(self.local_rep_of[order], self.local_reps)
        lo, hi = int(self.g_off[self.rank]), int(self.g_off[self.rank + 1])
        idx, dst, tag = [], [], []
        for c, (_r, _n, segs) in enumerate(sel):
            pos = 0
            for g, tk in segs:
                if lo <= g < hi:
                    s0 = int(starts[g - lo])
                    mem = order[s0:s0 + tk]
                    idx.append(mem); dst.append(np.full(tk, owner[c], dtype=np.int32))
                    tag.append((np.int64(c) << 32) | (pos + np.arange(tk, dtype=np.int64)))
                pos += tk
        idx = np.concatenate(idx) if idx else np.zeros(0, np.int64)
        dst = np.concatenate(dst) if dst else np.zeros(0, np.int32)
        tag = np.concatenate(tag) if tag else np.zeros(0, np.int64)
        srt = np.argsort(dst, kind="stable")
        expect = sum(u for u, o in zip(used, owner) if o == self.rank)
        tags, _cnt = eng.exchange_reads(idx[srt], dst[srt], tag[srt], ce, expect)
        t = self._tick("exchange_reads", t)
        mine = [c for c in range(len(sel)) if owner[c] == self.rank]
        by_tag = np.argsort(tags, kind="stable")
        lists, o = {}, 0
        cl_of = (tags[by_tag] >> 32).astype(np.int64)
        for c in mine:
            nread = used[c]
            assert (cl_of[o:o + nread] == c).all()
            lists[c] = by_tag[o:o + nread].tolist()
            o += nread
        ce.adopt_device_reads()
        drafts_mine, _nodes = C.draft_consensus_batch(ce, [lists[c] for c in mine]) if mine else ([], None)
        t = self._tick("draft", t)
        blobs = eng.allgather_bytes("\n".join(drafts_mine).encode())
        drafts = [None] * len(sel)
        for r, b in enumerate(blobs):
            got = b.decode().split("\n") if b else []
            for c, d in zip([c for c in range(len(sel)) if owner[c] == r], got):
                drafts[c] = d
        # ---- reverse-complement detection: every pair of drafts in both orientations, one K4 launch
        n_c = len(sel)
        ident = np.zeros((n_c, n_c))
        if n_c > 1:
            ii, jj = np.triu_indices(n_c, 1)
            aux = list(drafts) + [C.reverse_complement(d) for d in drafts]
            a = np.concatenate([-(ii + 1), -(ii + 1)]).astype(np.int32)
            b = np.concatenate([-(jj + 1), -(jj + n_c + 1)]).astype(np.int32)
            _s, m, cols = ce.sg_align_paths(a, b, np.full(len(a), 3, dtype=np.int32), aux=aux)
            idn = m / np.maximum(cols, 1).astype(np.float64)
            ident[ii, jj] = np.maximum(idn[:len(ii)], idn[len(ii):])
        finals = merge_reverse_complements([n for _r, n, _s in sel], ident, rc_identity_threshold)
        t = self._tick("rc_merge", t)
        # ---- reads of every surviving centre to its polishing owner
        f_used = [sum(used[c] for c in mem) for _i, _tot, mem in finals]
        f_owner = assign_owners(f_used, self.world)
        idx, dst, tag = [], [], []
        for f, (_i, _tot, mem) in enumerate(finals):
            pos = 0
            for c in mem:
                if owner[c] == self.rank:
                    li = np.asarray(lists[c], dtype=np.int64)
                    idx.append(li); dst.append(np.full(len(li), f_owner[f], dtype=np.int32))
                    tag.append((np.int64(f) << 32) | (pos + np.arange(len(li), dtype=np.int64)))
                pos += used[c]
        idx = np.concatenate(idx) if idx else np.zeros(0, np.int64)
        dst = np.concatenate(dst) if dst else np.zeros(0, np.int32)
        tag = np.concatenate(tag) if tag else np.zeros(0, np.int64)
        srt = np.argsort(dst, kind="stable")
        expect = sum(u for u, o2 in zip(f_used, f_owner) if o2 == self.rank)
        tags, _cnt = ce.exchange_reads(idx[srt], dst[srt], tag[srt], pe, expect)
        t = self._tick("exchange_reads", t)
        f_mine = [f for f in range(len(finals)) if f_owner[f] == self.rank]
        by_tag = np.argsort(tags, kind="stable")
        f_of = (tags[by_tag] >> 32).astype(np.int64)
        plists, o = [], 0
        for f in f_mine:
            assert (f_of[o:o + f_used[f]] == f).all()
            plists.append(by_tag[o:o + f_used[f]].tolist())
            o += f_used[f]
        n_fwd = pe.n_reads
        polished_mine = []
        if f_mine:
            pe.adopt_device_reads()
            pe.append_revcomp()
            rc_lists = [[n_fwd + i for i in li] for li in plists]
            polished_mine = C.polish_batch(pe, [drafts[finals[f][0]] for f in f_mine], plists, racon_iter, rc_lists)
        t = self._tick("polish", t)
        blobs = eng.allgather_bytes("\n".join(polished_mine).encode())
        polished = [None] * len(finals)
        for r, b in enumerate(blobs):
            got = b.decode().split("\n") if b else []
            for f, d in zip([f for f in range(len(finals)) if f_owner[f] == r], got):
                polished[f] = d
        self._tick("gather_consensus", t)
        centers = [[tot, int(self.g_gid[sel[i][0]]), polished[f]] for f, (i, tot, _m) in enumerate(finals)]
        info = {"clusters_selected": len(sel), "centres_after_rc_merge": len(finals),
                "reads_draft": int(sum(used)), "reads_polish": int(sum(f_used)),
                "reads_draft_this_rank": int(sum(used[c] for c in mine)),
                "reads_polish_this_rank": int(sum(f_used[f] for f in f_mine)),
                "drafts": drafts, "owner": owner}
        return centers, info
