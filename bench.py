#!/usr/bin/env python
"""bench.py -- reads/s of the per-read assignment hot path (banded Viterbi + SEP placement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--scale full|pe|long|small]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE for N > 1).  Workloads (BASELINE.json configs):
  full  = configs[2] (C3, the default and the headline): synthetic gg_97-scale database (99 322 leaves -> 198 643 nodes,
          7 682 consensus columns, GTR) + 250 bp V4-like SE reads;
  pe    = configs[3] (C4): the same database evaluated with GTR + dGamma4 (alpha 0.5) + 2 x 250 bp PE reads, overlapping mates;
  long  = configs[4] (C5): the same database with HKY85 + dGamma4 + ~1 400 bp near-full-length reads (windows of ~7 000 columns).
A "step" is one batch of reads through alignSeq -> getSeed -> estimateSeq -> filterPlacements -> placeSeq -> calcQValues
and the gather of the fixed-size assignment rows (NCCL all_gather on device buffers when N > 1).  Reads are sharded over
ranks, the database is replicated (weak scaling).

Rank 0 prints ONE JSON line.  See DESIGN.md "Measurement" for how each field is obtained.
"""
from __future__ import annotations

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

METRIC = "reads/s (banded Viterbi + SEP placement)"
GG97 = dict(n_leaves=99322, cs_len=7682, occupancy=0.18, seed=20260103)
SCALES = {
    # BASELINE.json configs[2]; window: all reads start inside one 300-column stretch
    "full": dict(GG97, window=(2000, 2300), read_len=250, model="GTR", dg=None, cpu_per_core=40,
                 workload="C3: synthetic gg_97-scale DB (99322 leaves, 198643 nodes, csLen 7682, GTR) + 250 bp V4-like SE reads"),
    # BASELINE.json configs[3]: the C3 database re-evaluated with dGamma K=4, alpha=0.5; 2 x 250 bp pairs, overlapping mates
    "pe": dict(GG97, window=None, read_len=250, paired=True, insert_len=420, model="GTR", dg=(0.5, 4), cpu_per_core=10,
               workload="C4: synthetic gg_97-scale DB (GTR + dGamma4, alpha 0.5) + 2x250 bp PE reads, overlapping mates (merged windows of 1.5-2.5 k columns)"),
    # BASELINE.json configs[4]: HKY85 + dGamma4, near-full-length reads (a leaf holds ~1 380 residues: N(1400, 50) clipped to it)
    "long": dict(GG97, window=(0, 300), read_len=1400, len_sd=50.0, model="HKY85", dg=(0.5, 4), cpu_per_core=2,
                 workload="C5: synthetic gg_97-scale DB (HKY85 + dGamma4, alpha 0.5) + ~1400 bp near-full-length SE reads (windows of ~7000 columns)"),
    # a shrunken stand-in for quick functional runs (NOT a benchmark configuration)
    "small": dict(n_leaves=2000, cs_len=2000, occupancy=0.5, seed=20260101, window=(300, 500), read_len=250, model="GTR", dg=None, cpu_per_core=40,
                  workload="reduced functional config (2000 leaves, csLen 2000) -- not a BASELINE workload"),
}


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu = gpu
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def build_inputs(cfg, n_reads, read_seed):
    """-> (db, (reads,) or (reads, mates)).  The tree and the sequences are those of C3 for every gg_97-scale shape; the
    substitution model and the discrete-Gamma rates the database is EVALUATED with are the shape's."""
    from hmmufotu_b200 import synth
    t0 = time.time()
    db = synth.make_db(n_leaves=cfg["n_leaves"], cs_len=cfg["cs_len"], occupancy=cfg["occupancy"], seed=cfg["seed"])
    db.model = synth.gg97_model(synth.MODEL_NAMES.index(cfg["model"]))
    if cfg.get("dg"):
        db.dg_rates = synth.discrete_gamma_rates(cfg["dg"][0], cfg["dg"][1])
    t1 = time.time()
    if cfg.get("paired"):
        sets = synth.make_pairs(db, n_reads, read_len=cfg["read_len"], insert_len=cfg["insert_len"], insert_sd=20.0, seed=read_seed)
    else:
        sets = (synth.make_reads(db, n_reads, read_len=cfg["read_len"], seed=read_seed, window=cfg["window"], len_sd=cfg.get("len_sd", 0.0)),)
    log("synthetic DB %.1fs (K=%d, nodes=%d, model %s%s), %d reads %.1fs" %
        (t1 - t0, db.K, db.n_nodes, cfg["model"], " + dGamma%d" % cfg["dg"][1] if cfg.get("dg") else "", n_reads, time.time() - t1))
    return db, sets


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    return O


class SlicedCpuReference:
    """The CPU oracle (reference-equivalent restatement; kind "port") over a read sample, on `threads` host threads, with
    no GPU anywhere.  The tree is cut to the consensus columns the reads can touch: every SEP quantity is per column, so the
    result is the one the full tree gives, and the oracle's 4 x csLen fp64 messages of all 397k directed edges (98 GB at
    full length) fit in host memory.  Building the cut tree (all-roots evaluation on the CPU) is untimed: it is the
    database load of the reference.  Only for shapes whose reads share a window (C3)."""

    def __init__(self, db, sets, threads):
        import copy
        O = self.O = _oracle()
        self.sets, self.threads = sets, threads
        self.ohmm = O.OracleHmm(db)
        lo, hi = db.cs_len, 0
        for x in sets:
            for i in range(x.n):  # alignment columns of the sample fix the slice
                a = self.ohmm.align(x.read_str(i), x.vpaths[i, :x.n_paths[i]])
                lo, hi = min(lo, int(a["coords"][4]) - 1), max(hi, int(a["coords"][5]) - 1)
        self.c0, self.c1 = max(0, lo - 8), min(db.cs_len, hi + 9)
        sl = copy.copy(db)
        sl.cs_len = self.c1 - self.c0
        sl.seq = np.ascontiguousarray(db.seq[:, self.c0:self.c1])
        t0 = time.time()
        self.otree = O.OracleTree(sl, evaluate=True, infer_internal=True, threads=threads)
        self.build_seconds = time.time() - t0
        self.how = "oracle tree cut to consensus columns %d..%d, evaluated on the CPU in %.0fs (untimed)" % (self.c0, self.c1, self.build_seconds)

    def run(self):
        t0 = time.time()
        rows, coords, cost, bad, ties = self.O.run_reads(self.ohmm, self.otree, self.sets[0], col0=self.c0, seed_order=0, threads=self.threads,
                                                         mates=self.sets[1] if len(self.sets) > 1 else None, with_ties=True)
        dt = time.time() - t0
        return dict(reads_per_s=self.sets[0].n / dt, seconds=dt, bad=bad, rows=rows, coords=coords, cost=cost, ties=ties, col0=self.c0)


class DeviceDbCpuReference:
    """The same CPU program over the SAME database the GPU arm holds: node sequences (inferred ones included) and the
    messages of the edges the sample can touch are read back from the device replica (hu_db_get_seq,
    hu_db_get_messages_of) -- the database is an input of the per-read path, as the .ptu file is for the reference -- so no
    98 GB of messages is ever built in host memory and the parity check compares the two arms on identical inputs.  The
    edges are the GPU's own seeds of the sample reads; a CPU seed outside that set is reported (row.c = -2) and counts as
    a parity mismatch.  Everything timed (alignSeq .. calcQValues for every read of the sample) runs on the host threads."""

    def __init__(self, db, gdb, sets, seed_ids, windows, threads):
        import copy
        O = self.O = _oracle()
        self.sets, self.threads = sets, threads
        self.ohmm = O.OracleHmm(db)
        t0 = time.time()
        ok = windows[:, 0] >= 0
        self.c0 = max(0, int(windows[ok, 0].min()) - 8) if ok.any() else 0
        self.c1 = min(db.cs_len, int(windows[ok, 1].max()) + 9) if ok.any() else db.cs_len
        seq = gdb.seqs()
        sl = copy.copy(db)
        sl.cs_len = self.c1 - self.c0
        sl.seq = np.ascontiguousarray(seq[:, self.c0:self.c1])
        del seq
        self.otree = O.OracleTree(sl, evaluate=False)
        ids = np.unique(seed_ids[seed_ids >= 0]).astype(np.int64)
        up, dn = gdb.messages_of(ids, self.c0, self.c1 - self.c0)
        self.otree.set_sparse_messages(ids, up, dn)
        self.build_seconds = time.time() - t0
        self.how = ("database read back from the device replica: all node sequences + the messages of %d edges (the sample's seeds), "
                    "consensus columns %d..%d, %.1f GB, %.0fs (untimed)" % (ids.size, self.c0, self.c1, 2 * up.nbytes / 1e9, self.build_seconds))

    run = SlicedCpuReference.run


def compare_rows(cpu, merged, best, n):
    """GPU rows (through the C ABI) against the CPU program's, read by read: alignment coordinates and cost bit-exact, the
    chosen branch / annotated node identical, ratio, wnr and log-likelihood to 1e-9 relative."""
    rows, coords, cost, ties = cpu["rows"], cpu["coords"], cpu["cost"], cpu["ties"]
    out = dict(checked=0, mismatched=0, near_tie_reordered=0, chimera=0, examples=[])

    def close(a, b):
        return a == b or abs(a - b) <= 1e-9 * max(abs(a), abs(b), 1e-300) or (a != a and b != b)

    for i in range(n):
        out["checked"] += 1
        r, g, m = rows[i], best[i], merged[i]
        bad = []
        if r.c == -3:                                   # PE orientation test failed on the CPU: the GPU must flag the same
            out["chimera"] += 1
            if not (m["flags"] & 8):
                bad.append("chimera flag")
        else:
            if [int(m[f]) for f in ("seqStart", "seqEnd", "hmmStart", "hmmEnd", "csStart", "csEnd")] != [int(v) for v in coords[i]]:
                bad.append("coords")
            if float(m["cost"]) != float(cost[i]):
                bad.append("cost %r != %r" % (float(m["cost"]), float(cost[i])))
            if r.c < 0:
                bad.append("cpu row invalid (%d)" % r.c)
            elif (int(g["cNode"]), int(g["pNode"])) != (r.c, r.p):
                if ties[i]:
                    out["near_tie_reordered"] += 1     # order of estimates that agree to 1e-11 is rounding noise in either program
                else:
                    bad.append("branch %d->%d != %d->%d" % (g["cNode"], g["pNode"], r.c, r.p))
            else:
                if int(g["aNode"]) != r.a:
                    bad.append("aNode")
                if int(g["nPlaces"]) != r.pad:
                    bad.append("nPlaces %d != %d" % (g["nPlaces"], r.pad))
                for f, v in (("ratio", r.ratio), ("wnr", r.wnr), ("loglik", r.loglik)):
                    if not close(float(g[f]), float(v)):
                        bad.append("%s %r != %r" % (f, float(g[f]), float(v)))
                for f, v in (("qPlace", r.qPlace), ("qTaxon", r.qTaxon)):   # -10 log10(1 - p): beyond Q ~ 150, 1 - p is rounding noise in either program
                    a, b_ = float(g[f]), float(v)
                    if not (a == b_ or abs(a - b_) <= 1e-7 * max(abs(a), abs(b_)) or (a != a and b_ != b_) or min(a, b_) > 150):
                        bad.append("%s %r != %r" % (f, a, b_))
        if bad:
            out["mismatched"] += 1
            if len(out["examples"]) < 4:
                out["examples"].append("read %d: %s" % (i, "; ".join(bad)))
    if not out["examples"]:
        del out["examples"]
    return out


def calibrate_reference(threads):
    """The reference's OWN binary (oracle/_ref/hmmufotu, built from /root/reference/src against the stand-in
    Eigen/Boost headers) and the oracle port on the same 128 reads of the reference-built 70_otus database
    (tests/golden/ref70, BASELINE.json configs[0]), both on `threads` host threads: how the port used for the
    gg_97-scale baselines relates to the real program.  None when the prebuilt binary is absent."""
    import gzip
    import shutil
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "hmmufotu")
    gold = os.path.join(ROOT, "tests", "golden", "ref70")
    name = "gg_70_otus_GTR"
    if not (os.path.exists(exe) and os.path.exists(os.path.join(gold, name + ".ptu.gz"))):
        return None
    try:
        O = _oracle()
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import refdb
        from types import SimpleNamespace
        d = tempfile.mkdtemp(prefix="hu_cal_")
        for ext in (".msa", ".hmm", ".csfm"):
            shutil.copy(os.path.join(gold, name + ext), os.path.join(d, name + ext))
        with gzip.open(os.path.join(gold, name + ".ptu.gz"), "rb") as f, open(os.path.join(d, name + ".ptu"), "wb") as g:
            shutil.copyfileobj(f, g)
        fa = os.path.join(d, "cal.fasta")
        with open(fa, "w") as out:
            for k in (0, 2):
                out.write(open(os.path.join(gold, "%s_se%d.fasta" % (name, k))).read())
        t0 = time.time()
        subprocess.run([exe, name, "cal.fasta", "-o", "cal.tsv", "-s", "1", "-S", "0", "-p", str(threads)], cwd=d, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        t_ref = time.time() - t0
        db = refdb.load_ptu(os.path.join(d, name + ".ptu"))
        tree = O.OracleTree(db, evaluate=False)
        up, dn = np.ascontiguousarray(db.up), np.ascontiguousarray(db.dn)
        O.lib().orc_tree_set_messages(tree.h, up.ctypes.data_as(O.C.c_void_p), dn.ctypes.data_as(O.C.c_void_p))
        hmm = O.OracleHmm(path=os.path.join(d, name + ".hmm"))
        reads = refdb.read_fasta(fa)
        seeds = subprocess.run([os.path.join(ROOT, "hmmufotu_b200", "bin", "hu_host_test"), "seeds", os.path.join(d, name), fa, "0"],
                               capture_output=True, text=True, check=True).stdout.strip().split("\n")
        vp = np.zeros((len(reads), 2, 6), dtype=np.int32); npth = np.zeros(len(reads), dtype=np.int32)
        for i, ln in enumerate(seeds):
            v = np.array([[int(x) for x in q.split(",")] for q in ln.split("\t")[2:]], dtype=np.int32).reshape(-1, 6)
            npth[i] = len(v); vp[i, :len(v)] = v
        codes = np.concatenate([np.array(["ACGT".index(c) for c in sq.upper()], dtype=np.int8) for _, _, sq in reads])
        offs = np.concatenate([[0], np.cumsum([len(sq) for _, _, sq in reads])]).astype(np.int64)
        rs = SimpleNamespace(n=len(reads), codes=codes, offsets=offs, vpaths=vp, n_paths=npth)
        t0 = time.time()
        O.run_reads(hmm, tree, rs, threads=threads)
        t_port = time.time() - t0
        shutil.rmtree(d, ignore_errors=True)
        return dict(db="70_otus GTR built by the reference's hmmufotu-build (configs[0])", reads=len(reads), threads=threads,
                    reference_binary_reads_per_s=len(reads) / t_ref, port_reads_per_s=len(reads) / t_port,
                    note="reference binary = oracle/_ref/hmmufotu (wall time incl. its 0.2 s database load)")
    except Exception as e:  # a calibration failure must not take the bench line down
        return dict(error=str(e)[:200])


def main():
    # stdout carries exactly ONE JSON line: everything else that libraries print there (e.g. NCCL's version
    # banner) is sent to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rc = _main(real_stdout)
    real_stdout.flush()
    return rc


def make_config(cfg, B, world, n_ws, db_gb=None, n_nodes=None):
    """`config` of the JSON line -- the same dict for both arms (the reference arm runs a bounded sample of this workload)."""
    c = dict(workload=cfg["workload"], batch=B, parallelism="replicated DB, reads sharded x%d, rows gathered by one NCCL all_gather per batch" % world,
             l2="inputs larger than L2 (DB ~126 GB per GPU, per-batch key matrix %.1f GB)" % (B * (2 * cfg["n_leaves"] - 1) * 4 / 1e9),
             workspaces=n_ws)
    return c


def gpu_sample_seeds(engine, gdb, sets, sample, max_len, paired):
    """The GPU's seeds and windows of the first `sample` reads (what decides which edges the CPU arm's sparse database needs),
    and its rows for the parity check."""
    b = engine.Batch(gdb, sample, max_len, paired=paired)
    b.upload(*[x.slice(0, sample) for x in sets])
    b.assign(engine.SepParams())
    ns, ids, _ = b.seeds()
    merged = b.merged()
    best = b.best()
    win = np.stack([merged["csStart"] - 1, merged["csEnd"] - 1], axis=1).astype(np.int64)
    win[(merged["flags"] & (8 | 16)) != 0] = -1
    heap = b.stats()["seedHeapSorts"]
    b.close()
    return ids, win, merged, best, heap


def _main(real_stdout):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", default=os.environ.get("HU_BENCH_SCALE", "full"), choices=sorted(SCALES))
    ap.add_argument("--batch", type=int, default=int(os.environ.get("HU_BENCH_BATCH", "0")))
    ap.add_argument("--cpu-sample", type=int, default=int(os.environ.get("HU_BENCH_CPU_SAMPLE", "0")))
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workspaces", type=int, default=int(os.environ.get("HU_BENCH_WORKSPACES", "2")))
    args = ap.parse_args()
    cfg = SCALES[args.scale]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    paired = bool(cfg.get("paired"))
    B = args.batch or (2048 if args.scale == "long" else 8192)
    n_ws = max(1, args.workspaces)
    sample = args.cpu_sample or max(cores * cfg["cpu_per_core"], 32)

    if args.impl == "reference":
        # the reference arm: the CPU program on all host threads over a bounded sample of the same workload; rank 0 only
        if rank != 0:
            return 0
        db, sets = build_inputs(cfg, sample, read_seed=1000)
        if args.scale in ("full", "small"):
            ref = SlicedCpuReference(db, sets, cores)                  # no GPU anywhere in this arm
        else:
            # reads of these shapes touch the whole consensus: the database (all-roots evaluation, 98 GB of messages) is
            # produced on the device as for the product arm and the needed part read back; the timed region is CPU only
            import torch
            torch.cuda.set_device(local)
            from hmmufotu_b200 import engine
            gdb = engine.Database(db, device=local)
            max_len = max(int(np.max(np.diff(x.offsets))) for x in sets) + 8
            ids, win, _, _, _ = gpu_sample_seeds(engine, gdb, sets, sample, max_len, paired)
            ref = DeviceDbCpuReference(db, gdb, sets, ids, win, cores)
            gdb.close()
        vals = []
        for step in range(args.warmup + args.steps):
            r = ref.run()
            if step >= args.warmup:
                vals.append(r)
            log("reference step %d: %.2f reads/s (%d reads, %.1fs; %s)" % (step, r["reads_per_s"], sample, r["seconds"], ref.how))
        tot = sum(v["seconds"] for v in vals)
        v = sample * len(vals) / tot
        line = dict(metric=METRIC, value=v, unit="reads/s", impl="reference", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                    ms_per_step=1e3 * tot / len(vals), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                    data="synthetic", config=make_config(cfg, B, args.gpus, n_ws),
                    cpu_baseline=dict(value=v, unit="reads/s", cores=cores, kind="port",
                                      sample="%d reads per step on %d threads; %s" % (sample, cores, ref.how),
                                      calibration=calibrate_reference(cores)),
                    e2e=dict(value=v, unit="reads/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
        print(json.dumps(line), file=real_stdout, flush=True)
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from hmmufotu_b200 import driver, engine

    pool_batches = 4
    db, sets = build_inputs(cfg, B * pool_batches, read_seed=1 + rank)
    t0 = time.time()
    gdb = engine.Database(db, device=local)          # uploads, evaluates all messages on the device, repacks
    log("rank %d: device DB %.1f GB in %.1fs" % (rank, gdb.bytes / 1e9, time.time() - t0))
    max_len = max(int(np.max(np.diff(x.offsets))) for x in sets) + 8
    ws = [engine.Batch(gdb, B, max_len, paired=paired) for _ in range(n_ws)]
    gathers = [driver.DeviceGather(B, world, local) for _ in range(n_ws)]
    params = engine.SepParams()
    batches = []
    for i in range(pool_batches):
        parts = []
        for x in sets:
            s = x.slice(i * B, (i + 1) * B)
            s.codes = engine.pinned_copy(s.codes); s.offsets = engine.pinned_copy(s.offsets)
            s.vpaths = engine.pinned_copy(s.vpaths); s.n_paths = engine.pinned_copy(s.n_paths)
            parts.append(s)
        batches.append(parts)

    def barrier():
        for w in ws:
            w.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_resident(i):
        """one step with the inputs already in HBM: the five stages + the rows' gather, all enqueued, nothing waited for"""
        k = i % n_ws
        ws[k].assign(params)
        gathers[k].launch(ws[k], (rank * args.steps + i) * B)

    def drain():
        for k in range(n_ws):
            gathers[k].finish_on_stream()           # the batch stream waits for its gather, so that the end mark covers it

    # ---------------- kernel-resident throughput: inputs already in HBM ----------------
    for i, w in enumerate(ws):
        w.upload(*batches[i % pool_batches])
    for i in range(max(args.warmup, 1)):
        step_resident(i)
    drain()
    barrier()
    clk = ClockSampler(local)
    clk.start()
    stage = np.zeros(8)
    for w in ws:
        w.mark(0)
    for step in range(args.steps):
        step_resident(step)
    drain()
    for w in ws:
        w.mark(1)
    ms = max(ws[0].elapsed_ms(0, w, 1) for w in ws)
    barrier()
    clocks = clk.stop()
    ms = max_over_ranks(ms)
    value = world * B * args.steps / (ms / 1e3)
    # per-stage device times (CUDA events on the launching stream) and work counters: the same step run
    # alone on one workspace, so that a stage's span is not stretched by the other stream's kernels
    stats = {}
    n_meas = 2
    for rep in range(n_meas):
        ws[0].assign(params)
        ws[0].sync()
        stage += ws[0].stage_ms()
        for k, v in ws[0].stats().items():
            stats[k] = stats.get(k, 0) + v
    launches_per_step = stats["launches"] / n_meas + 1          # + the row-packing kernel of the gather
    if os.environ.get("HU_PROF"):  # a library built with HU_PROF=1 carries phase clocks of CTA 0 (see hu_seed.cu, hu_estimate.cu)
        import ctypes
        from hmmufotu_b200 import _lib
        L = _lib.load()
        for sym in ("hu_prof_seed", "hu_prof_est"):
            if hasattr(L, sym):
                buf = (ctypes.c_ulonglong * 16)()
                getattr(L, sym)(buf)
                log("%s %s" % (sym, list(buf)))

    # ---------------- end to end: host buffers in, assignment rows + traces out, every step ----------------
    d2h_step = [0]

    def e2e_step(i):
        w = ws[i % n_ws]
        w.upload(*batches[i % pool_batches])
        w.assign(params)
        gathers[i % n_ws].launch(w, (rank * args.steps + i) * B)

    def e2e_collect(i):
        w = ws[i % n_ws]
        rows, (tbuf, toff) = w.alignments(with_traces="packed")    # alignment rows + packed traces into page-locked buffers
        if world > 1:
            g = gathers[i % n_ws].rows() if rank == 0 else (gathers[i % n_ws].wait(), None)[1]   # gathered rows to the host on rank 0
            d2h_step[0] = rows.nbytes + int(toff[-1]) + toff.nbytes + (g.nbytes if g is not None else 0)
        else:
            best = w.best(pinned=True)
            d2h_step[0] = rows.nbytes + int(toff[-1]) + toff.nbytes + best.nbytes
        return rows

    for i in range(2):
        e2e_step(i)
        e2e_collect(i)
    barrier()
    t0 = time.perf_counter()
    # two workspaces: the read-back of step i overlaps the kernels of step i+1
    e2e_step(0)
    for i in range(args.steps):
        if i + 1 < args.steps:
            e2e_step(i + 1)
        e2e_collect(i)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * B * args.steps / e2e_s
    h2d = int(sum(x.codes.nbytes for x in batches[0]) + (B * 312 * len(batches[0])))   # base codes + band plans (312 B per sequence)
    d2h = int(d2h_step[0])

    # ---------------- roofline of every stage, and of the dominant one ----------------
    names = ["viterbi (K1)", "seed scan (K2)", "seed select (K2b)", "estimate (K3)", "place (K4)", "finalize"]
    per = stage[:6] / n_meas
    dom = int(np.argmax(per))
    peaks = dict(hbm_gbs=6650.0, src="fallback")
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peaks = dict(hbm_gbs=float(pk["hbm_gbs"]), src="measured")
    except Exception:
        pass
    # DRAM bytes per launch of each stage's kernels from the last committed `ncu --set full` capture (C3 shape only)
    traffic = {}
    if args.scale == "full":
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            pass
    sm_clock = (clocks.get("sm_mhz") or 1965.0) * 1e6
    n_sm = torch.cuda.get_device_properties(local).multi_processor_count
    # algorithmic work per step: bytes for the streaming stages, cell updates for Viterbi;
    # the literal seed order reads each read's key vector once (4 B per candidate node)
    alg = {1: stats["seedBytes"] / n_meas, 2: 4.0 * (db.n_nodes - 1) * B, 3: stats["estBytes"] / n_meas, 4: stats["placeBytes"] / n_meas}
    stages = {}
    for i, nme in enumerate(names):
        d = dict(ms=float(per[i]))
        if i in alg and per[i] > 0:
            d["algorithmic_GBps"] = alg[i] / (per[i] / 1e3) / 1e9
            d["hbm_frac"] = d["algorithmic_GBps"] / peaks["hbm_gbs"]
        if i == 0 and per[0] > 0:
            d["gcups"] = stats["cells"] / n_meas / (per[0] / 1e3) / 1e9
            # 15 ALU operations per cell update against one issue per lane and clock (fp64 add/min: SURVEY.md 8(d) K1)
            d["alu_issue_frac"] = d["gcups"] * 1e9 * 15 / (n_sm * 128 * sm_clock)
        if i == 1 and per[1] > 0 and stats.get("seedMacs", 0) > 0:
            # the scan runs on the tensor cores (tcgen05.mma kind::i8): multiply-accumulates of the tiles x chunks visited;
            # nominal dense int8 peak of a B200 (no measured int8 figure in MEASURED_PEAKS.json)
            d["int8_tops"] = 2.0 * stats["seedMacs"] / n_meas / (per[1] / 1e3) / 1e12
            d["int8_frac_of_nominal_4500"] = d["int8_tops"] / 4500.0
            d.pop("hbm_frac", None)   # SURVEY's byte count assumes one node byte per (read, node, column): not what bounds this kernel
        if nme in traffic:
            d["dram_bytes_per_launch"] = traffic[nme]
        stages[nme] = d
    achieved = alg.get(dom, 0.0) / (per[dom] / 1e3) / 1e9 if per[dom] > 0 else 0.0
    roofline = dict(bound="hbm", kernel=names[dom], achieved=achieved, peak=peaks["hbm_gbs"], unit="GB/s",
                    frac=achieved / peaks["hbm_gbs"], traffic=traffic.get(names[dom]), peak_source=peaks["src"],
                    note="achieved = algorithmic bytes per step (SURVEY.md 8(d)) / CUDA-event time of the stage; the seed scan is an "
                         "int8 contraction on the tensor cores (tcgen05.mma), reported as int8_tops under 'stages', where every "
                         "stage has its own figures")

    if dom == 0 and stages[names[0]].get("gcups"):   # long reads: the Viterbi recurrence dominates; its roofline is ALU issue
        g = stages[names[0]]
        roofline = dict(bound="alu", kernel=names[0], achieved=g["gcups"] * 15, peak=n_sm * 128 * sm_clock / 1e9, unit="Gop/s",
                        frac=g["alu_issue_frac"], traffic=traffic.get(names[0]), peak_source="148 SMs x 128 lanes x measured SM clock",
                        note="15 ALU operations per cell update (10 fp64 adds, 5 compares) against one issue per lane and clock")

    # ---------------- CPU baseline + parity of the two arms on the same reads (rank 0, N = 1) ----------------
    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = min(sample, B)
        sub = [x.slice(0, sample) for x in sets]
        ids, win, merged, best, heap = gpu_sample_seeds(engine, gdb, sets, sample, max_len, paired)
        ref = DeviceDbCpuReference(db, gdb, sub, ids, win, cores)
        r = ref.run()
        parity = compare_rows(r, merged, best, sample)
        parity["seed_heap_sorts"] = int(heap)
        parity["note"] = ("GPU rows through the C ABI vs the CPU program (oracle port, pinned to the reference binary in tests/test_ref_pin.py) on the same %d reads and "
                          "the same database: coords + cost bit-exact, branch / node identical, ratio, wnr, loglik, Q values to 1e-9" % sample)
        log("parity: %s" % json.dumps(parity))
        # kind "port": the reference's own binary (oracle/_ref) needs the database as files, and a gg_97-scale .ptu is
        # 98 GB; the port is pinned to that binary row by row (tests/test_ref_pin.py) and is the FASTER CPU program on the
        # 70_otus database, where both run (see `calibration`), so it is the conservative baseline
        cpu = dict(value=r["reads_per_s"], unit="reads/s", cores=cores, kind="port",
                   sample="%d reads of the same pool, %d threads, %.1fs; %s" % (sample, cores, r["seconds"], ref.how),
                   calibration=calibrate_reference(cores) if args.scale == "full" else None)

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return 0
    line = dict(metric=METRIC, value=value, unit="reads/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=make_config(cfg, B, world, n_ws),
                e2e=dict(value=e2e_value, unit="reads/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                gpu_launches=int(launches_per_step * args.steps), roofline=roofline, cpu_baseline=cpu, parity=parity, clocks=clocks,
                viterbi_gcups=stages[names[0]].get("gcups"), stages=stages)
    print(json.dumps(line), file=real_stdout, flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
