"""
Reference arm of bench.py (`bench.py --impl reference`) and its `cpu_baseline` leg: the UNMODIFIED
reference (ksahlin/NGSpeciesID v0.3.1, installed once into baseline/_ref with
`pip install --no-index --no-deps --target baseline/_ref /root/reference`, git-ignored) clustering
a bounded sample of the bench workload on the host cores.

What runs is the reference's own code: modules/cluster.py:reads_to_clusters through the call the
NGSpeciesID script makes for --t 1 (NGSpeciesID:20-33) or modules/parallelize.py:parallel_clustering
for --t N, on the read_array the script builds (NGSpeciesID:58), with the probability table of
modules/p_minimizers_shared.py. The only stand-in is the `parasail` module (third party, not in this
image): oracle/parasail_shim drives oracle/sg_align.c; the seconds spent inside it are reported.
When baseline/_ref is missing the oracle port (oracle/cluster_oracle.py) is timed instead and the
line says kind = "port".

Replicas: `--t T` uses T cores at most; to use the whole host, cores // T independent copies of the
same job run at the same time (separate processes, started together); the reported rate is
copies x sample / slowest copy.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
SHIM = os.path.join(ROOT, "oracle", "parasail_shim")


def have_reference():
    return os.path.isfile(os.path.join(REF, "modules", "cluster.py"))


def reference_args(k, w, t, outfolder):
    """The knobs of the clustering path at the reference's CLI defaults (NGSpeciesID:188-245)."""
    return argparse.Namespace(k=k, w=w, min_shared=5, mapped_threshold=0.7, aligned_threshold=0.4,
                              symmetric_map_align_thresholds=False, batch_type="total_nt", min_fraction=0.8,
                              min_prob_no_hits=0.1, nr_cores=t, outfolder=outfolder, print_output=10000)


def _load_sample(path):
    z = np.load(path, allow_pickle=False)
    seq, qual, off = z["seq"], z["qual"], z["offsets"]
    acc = [a.decode() for a in z["acc"]]
    ra = []
    for i in range(len(acc)):
        a, b = off[i], off[i + 1]
        ra.append((i, 0, acc[i], seq[a:b].tobytes().decode(), qual[a:b].tobytes().decode(), float(acc[i].split("_")[-1])))
    return ra


def worker_main(argv):
    """One copy of the job: `python reference_arm.py worker <sample.npz> <k> <w> <t> <iters> <start_epoch> <kind>`.
    Prints one JSON line: per-iteration seconds, seconds inside the aligner, clusters found."""
    sample, k, w, t, iters, start_at, kind = argv[0], int(argv[1]), int(argv[2]), int(argv[3]), int(argv[4]), float(argv[5]), argv[6]
    import tempfile
    out = tempfile.mkdtemp(prefix="ngsid_ref_")
    ra = _load_sample(sample)
    if kind == "reference":
        sys.path[:0] = [SHIM, REF]
        os.environ["PYTHONPATH"] = os.pathsep.join([SHIM, REF, os.environ.get("PYTHONPATH", "")])   # spawned pool workers of --t N
        import parasail                                   # the shim
        from modules import cluster, parallelize, p_minimizers_shared
        p_emp = {}
        for kk, ww, p, e1, e2 in p_minimizers_shared.read_empirical_p():      # NGSpeciesID:72-77
            if int(kk) == k and abs(int(ww) - w) <= 2:
                p_emp[(float(e1), float(e2))] = float(p)
                p_emp[(float(e2), float(e1))] = float(p)
        args = reference_args(k, w, t, out)

        def run():
            if t > 1:
                return parallelize.parallel_clustering(list(ra), p_emp, args)
            clusters = {r[0]: [r[2]] for r in ra}
            reps = {r[0]: tuple(r) for r in ra}
            res = cluster.reads_to_clusters(clusters, reps, ra, p_emp, {}, 1, args)
            return list(res.values())[0][:2]
        calls = parasail.CALLS
    else:
        sys.path.insert(0, ROOT)
        from oracle import cluster_oracle as oc
        from ngspeciesid_b200.modules import p_minimizers_shared
        p_emp = p_minimizers_shared.p_emp_for(k, w)
        args = oc.default_args(nr_cores=t)
        calls = {"n": 0, "seconds": 0.0}

        def run():
            return oc.parallel_clustering(list(ra), p_emp, args) if t > 1 else oc.single_clustering(ra, p_emp, args)
    while time.time() < start_at:
        time.sleep(0.005)
    secs = []
    n_clusters = 0
    for _ in range(iters):
        t0 = time.perf_counter()
        cl, _rp = run()
        secs.append(time.perf_counter() - t0)
        n_clusters = len(cl)
    print(json.dumps({"seconds": secs, "aligner_seconds": calls.get("seconds", 0.0), "aligner_calls": calls.get("n", 0),
                      "clusters": n_clusters, "reads": len(ra)}))


def write_sample(path, seq, qual, offsets, acc, n):
    tmp = "%s.%d.tmp.npz" % (path, os.getpid())
    np.savez(tmp, seq=seq[:offsets[n]], qual=qual[:offsets[n]], offsets=offsets[:n + 1],
             acc=np.array([a.encode() for a in acc[:n]]))
    os.replace(tmp, path)


def run_copies(sample_path, k, w, t, iters, copies, kind):
    """Starts `copies` workers together; returns their parsed JSON lines."""
    start_at = time.time() + 4.0 + 0.05 * copies             # imports of every worker are done by then
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "worker", sample_path, str(k), str(w), str(t),
                               str(iters), repr(start_at), kind], stdout=subprocess.PIPE, text=True)
             for _ in range(copies)]
    outs = []
    for p in procs:
        o, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("reference worker failed")
        outs.append(json.loads(o.strip().splitlines()[-1]))
    return outs


def measure(seq, qual, offsets, acc, k, w, t, steps, warmup, cores, budget_s=150.0, max_reads=50000, tag="ref"):
    """Rate of the reference clustering `--t t` on a prefix of the score-ordered workload, sized so
    that warmup + steps iterations take about `budget_s` seconds. Returns a dict for the JSON line."""
    kind = "reference" if have_reference() else "port"
    copies = max(1, cores // max(1, t))
    base = "/tmp/ngsid_refsample_%s_%d" % (tag, os.getpid())
    # calibration: one copy, a few hundred reads per batch
    n_cal = min(len(acc), 400 * max(1, t))
    write_sample(base + "_cal.npz", seq, qual, offsets, acc, n_cal)
    cal = run_copies(base + "_cal.npz", k, w, t, 1, 1, kind)[0]
    rate1 = n_cal / max(1e-6, cal["seconds"][0])
    per_iter = budget_s / max(1, steps + warmup)
    ns = int(max(min(len(acc), 300 * max(1, t)), min(len(acc), max_reads, rate1 * per_iter)))
    write_sample(base + ".npz", seq, qual, offsets, acc, ns)
    outs = run_copies(base + ".npz", k, w, t, steps + warmup, copies, kind)
    per_step = [max(o["seconds"][i] for o in outs) for i in range(warmup, warmup + steps)]
    total = sum(per_step)
    for f in (base + "_cal.npz", base + ".npz"):
        try:
            os.remove(f)
        except OSError:
            pass
    single = outs[0]
    return {
        "kind": kind, "value": copies * ns * steps / total, "ms_per_step": 1000.0 * total / steps,
        "cores": min(cores, copies * max(1, t)), "copies": copies, "t": t, "sample_reads": ns,
        "one_copy_reads_per_s": ns * steps / sum(single["seconds"][warmup:]),
        "aligner_seconds_share": (single["aligner_seconds"] / sum(single["seconds"])) if t == 1 and sum(single["seconds"]) > 0 else None,
        "aligner_calls_per_iter": single["aligner_calls"] / float(steps + warmup) if t == 1 else None,
        "clusters": single["clusters"],
        "sample": ("prefix of %d reads of the same score-ordered workload, %d concurrent cop%s of the %s's --t %d clustering "
                   "(%s), aligner = oracle/sg_align.c behind the parasail shim"
                   % (ns, copies, "y" if copies == 1 else "ies",
                      "unmodified reference" if kind == "reference" else "oracle port", t,
                      "modules/cluster.py:reads_to_clusters via NGSpeciesID:20-33" if t == 1 else "modules/parallelize.py:parallel_clustering")),
    }


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "worker":
        worker_main(sys.argv[2:])
