"""-m gpu: wider parity coverage -- every model family with discrete Gamma, NGCL mode, paired-end merge and
orientation test, option filters, wide windows (global-memory slabs), inserts in non-match columns."""
import numpy as np
import pytest

import oracle_py as O
from hmmufotu_b200 import synth
from test_gpu_parity import RTOL, TIE_RTOL, _check_read

pytestmark = pytest.mark.gpu


def _setup(engine, **kw):
    db = synth.make_db(**kw)
    otree = O.OracleTree(db)
    up, dn = otree.messages()
    db.seq = otree.seqs()
    return db, otree, engine.Database(db, messages=(up, dn))


def _oracle_aligned(ohmm, rs, L, n):
    codes = np.full((n, L), -2, dtype=np.int8)
    start = np.zeros(n, dtype=np.int32); end = np.zeros(n, dtype=np.int32)
    for i in range(n):
        o = ohmm.align(rs.read_str(i), rs.vpaths[i, :rs.n_paths[i]])
        codes[i] = O.digitize(o["align"])
        start[i], end[i] = o["coords"][4] - 1, o["coords"][5] - 1
    return codes, start, end


@pytest.mark.parametrize("mtype,alpha", [(synth.GTR, 0.6), (synth.TN93, 0.9), (synth.HKY85, 0.0), (synth.F81, 1.5), (synth.K80, 0.4), (synth.JC69, 0.0)])
def test_models_and_discrete_gamma(mtype, alpha):
    from hmmufotu_b200 import engine
    db, otree, gdb = _setup(engine, n_leaves=30, cs_len=420, occupancy=0.85, seed=100 + mtype, model=synth.gg97_model(mtype), dg_alpha=alpha)
    # messages evaluated on the device agree with the oracle's
    g2 = engine.Database(db, messages=None, infer_internal=False)
    up, dn = g2.messages()
    oup, odn = otree.messages()
    nz = np.arange(db.n_nodes) != db.root
    for a, b in ((up[nz], oup[nz]), (dn[nz], odn[nz])):
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin)
        np.testing.assert_allclose(a[fin], b[fin], rtol=1e-11)
    g2.close()
    rs = synth.make_reads(db, 24, read_len=150, seed=9)
    ohmm = O.OracleHmm(db)
    n = rs.n
    codes, start, end = _oracle_aligned(ohmm, rs, db.cs_len, n)
    b = engine.Batch(gdb, n, 200)
    b.set_aligned(codes, start, end)
    b.sep(engine.SepParams())
    ns, ids, dist = b.seeds(with_dist=True)
    nk, est = b.estimates()
    _, pl = b.places()
    best = b.best()
    for i in range(n):
        _check_read(otree, codes[i], start[i], end[i], ns[i], ids[i], dist[i], nk[i], est[i], pl[i], best[i])
    b.close(); gdb.close()


def test_ngcl_mode_bit_exact():
    """--single: NGCL, every row of S can hold the minimum (T_SP(C,C) > 0)."""
    from hmmufotu_b200 import engine
    db, otree, gdb = _setup(engine, n_leaves=30, cs_len=500, occupancy=0.85, seed=31)
    rs = synth.make_reads(db, 80, read_len=160, seed=2, drop_seed_prob=0.1)
    ohmm = O.OracleHmm(db, mode=O.NGCL)
    b = engine.Batch(gdb, rs.n, 200)
    b.upload(rs)
    b.alignSeq(mode=2)
    rows, traces = b.alignments()
    for i in range(rs.n):
        o = ohmm.align(rs.read_str(i), rs.vpaths[i, :rs.n_paths[i]])
        assert [rows[i][k] for k in ("seqStart", "seqEnd", "hmmStart", "hmmEnd", "csStart", "csEnd")] == list(o["coords"]), i
        assert rows[i]["cost"] == o["cost"] and traces[i] == o["trace"], i
        assert gdb.trace_to_align(rs.read_str(i), rows[i], traces[i]) == o["align"], i
    b.close(); gdb.close()


def test_paired_end_merge_and_orientation():
    from hmmufotu_b200 import engine, _lib
    db, otree, gdb = _setup(engine, n_leaves=40, cs_len=700, occupancy=0.85, seed=41)
    fwd, rev = synth.make_pairs(db, 60, read_len=120, insert_len=230, seed=3, swap_prob=0.15)
    ohmm = O.OracleHmm(db)
    b = engine.Batch(gdb, fwd.n, 160, paired=True)
    b.upload(fwd, rev)
    b.assign(engine.SepParams())
    rows, traces = b.alignments()
    merged = b.merged()
    codes = b.aligned_codes()
    best = b.best()
    n = fwd.n
    n_chim = 0
    for i in range(n):
        a = ohmm.align(fwd.read_str(i), fwd.vpaths[i, :fwd.n_paths[i]])
        m = ohmm.align(rev.read_str(i), rev.vpaths[i, :rev.n_paths[i]])
        assert rows[i]["cost"] == a["cost"] and traces[i] == a["trace"]
        assert rows[n + i]["cost"] == m["cost"] and traces[n + i] == m["trace"]
        chim = not (a["coords"][4] <= m["coords"][4] and a["coords"][5] <= m["coords"][5])   # src/hmmufotu.cpp:629
        assert bool(merged[i]["flags"] & _lib.ALIGN_CHIMERA) == chim, i
        n_chim += chim
        if chim:
            assert best[i]["nPlaces"] == 0          # not placed; the reference writes it to --chimera-out only
            continue
        mm = O.merge(db.cs_len, a, m)
        assert [merged[i][k] for k in ("seqStart", "seqEnd", "hmmStart", "hmmEnd", "csStart", "csEnd")] == list(mm["coords"])
        assert merged[i]["cost"] == mm["cost"]
        want = O.digitize(mm["align"])
        got = codes[i].copy(); got[got < 0] = -2
        s, e = mm["coords"][4] - 1, mm["coords"][5] - 1
        assert np.array_equal(got[s:e + 1], want[s:e + 1]), i
        ids, dist = otree.get_seed(want, s, e, 50, order=0)
        ref = otree.assign(want, s, e, ids, dist)
        est = sorted(w.estLoglik for w in ref)
        if not any(abs(x - y) <= TIE_RTOL * abs(x) for x, y in zip(est, est[1:])):
            assert (best[i]["cNode"], best[i]["pNode"], best[i]["aNode"]) == (ref[0].c, ref[0].p, ref[0].a), i
            np.testing.assert_allclose([best[i]["ratio"], best[i]["wnr"]], [ref[0].ratio, ref[0].wnr], rtol=RTOL, atol=1e-12)
    assert 0 < n_chim < n
    b.close(); gdb.close()


def test_option_filters():
    """-N (maxNSeed), -e (maxError), -H (maxHeight), -d (maxDiff)."""
    from hmmufotu_b200 import engine
    db, otree, gdb = _setup(engine, n_leaves=50, cs_len=500, occupancy=0.85, seed=51)
    rs = synth.make_reads(db, 24, read_len=150, seed=6)
    ohmm = O.OracleHmm(db)
    codes, start, end = _oracle_aligned(ohmm, rs, db.cs_len, rs.n)
    b = engine.Batch(gdb, rs.n, 200)
    b.set_aligned(codes, start, end)
    hmax = float(np.median(db.height))
    p = engine.SepParams(max_n_seed=12, max_error=3.0, max_height=hmax)
    b.sep(p)
    ns, ids, dist = b.seeds(with_dist=True)
    nk, est = b.estimates()
    for i in range(rs.n):
        # getSeed filters by height first and std::sorts what is left (src/HmmUFOtu_main.cpp:131-139): literal order on the compacted vector
        oid, odist = otree.get_seed(codes[i], start[i], end[i], 12, max_height=hmax, order=0)
        assert np.array_equal(ids[i, :ns[i]], oid) and np.array_equal(dist[i, :ns[i]], odist)
        assert (db.height[oid] <= hmax).all()
        rows = otree.assign(codes[i], start[i], end[i], oid, odist, max_error=3.0, max_height=hmax)
        assert nk[i] == len(rows)
        assert est[i, 0]["estLoglik"] - est[i, nk[i] - 1]["estLoglik"] <= 3.0 + 1e-9
    # maxDiff: seeds further than d from the best are cut when the overall spread allows it
    b.getSeed(engine.SepParams(max_diff=0.02))
    ns2, ids2, _ = b.seeds()
    for i in range(rs.n):
        oid, odist = otree.get_seed(codes[i], start[i], end[i], 50, max_diff=0.02, order=0)
        assert np.array_equal(ids2[i, :ns2[i]], oid), i
    b.close(); gdb.close()


def test_wide_windows_and_inserts():
    """Windows wider than the shared-memory capacities of K4 (1536 columns) and reads with many insert bases
    placed in non-match columns (the sparse path of the seed scan)."""
    from hmmufotu_b200 import engine
    db, otree, gdb = _setup(engine, n_leaves=24, cs_len=4000, occupancy=0.22, seed=61, leaf_ins=0.004)
    rs = synth.make_reads(db, 16, read_len=420, seed=8, indel_prob=0.9)
    ohmm = O.OracleHmm(db)
    b = engine.Batch(gdb, rs.n, 480)
    b.upload(rs)
    b.assign(engine.SepParams())
    rows, traces = b.alignments()
    codes = b.aligned_codes()
    ns, ids, dist = b.seeds(with_dist=True)
    nk, est = b.estimates()
    _, pl = b.places()
    best = b.best()
    widths = []
    n_sparse = 0
    for i in range(rs.n):
        o = ohmm.align(rs.read_str(i), rs.vpaths[i, :rs.n_paths[i]])
        assert rows[i]["cost"] == o["cost"] and traces[i] == o["trace"]
        seq = O.digitize(o["align"])
        s, e = o["coords"][4] - 1, o["coords"][5] - 1
        widths.append(e - s + 1)
        n_sparse += int(((seq[s:e + 1] >= 0) & (db.cs2p[s + 1:e + 2] == 0)).sum())
        _check_read(otree, seq, s, e, ns[i], ids[i], dist[i], nk[i], est[i], pl[i], best[i])
    assert max(widths) > 1536 and n_sparse > 0
    b.close(); gdb.close()


def test_literal_seed_order_on_a_larger_tree():
    """~4 000 candidates per read: several levels of the parallel Hoare partition before ranges get small."""
    from hmmufotu_b200 import engine, _lib
    db = synth.make_db(n_leaves=2000, cs_len=260, occupancy=0.9, seed=71, mean_blen=0.02)
    otree = O.OracleTree(db, threads=8)
    db.seq = otree.seqs()
    gdb = engine.Database(db, messages=otree.messages())
    rs = synth.make_reads(db, 48, read_len=90, seed=12, drop_seed_prob=0.0)
    ohmm = O.OracleHmm(db)
    codes, start, end = _oracle_aligned(ohmm, rs, db.cs_len, rs.n)
    b = engine.Batch(gdb, rs.n, 128)
    b.set_aligned(codes, start, end)
    n_tie_cut = 0
    for max_n in (50, 17, 1):
        b.getSeed(engine.SepParams(max_n_seed=max_n, seed_order=_lib.SEED_ORDER_STDSORT))
        ns, ids, _ = b.seeds()
        for i in range(rs.n):
            oid, odist = otree.get_seed(codes[i], start[i], end[i], max_n, order=0)
            assert np.array_equal(ids[i, :ns[i]], oid), (max_n, i)
            sid, _ = otree.get_seed(codes[i], start[i], end[i], max_n, order=1)
            n_tie_cut += set(sid) != set(oid)
    assert n_tie_cut > 0   # tie groups cut by the truncation: std::sort and (dist, id) keep different members
    b.close(); gdb.close()


def test_long_reads_hky85_dgamma_wide_kernels():
    """BASELINE.json configs[4]'s shape on a mid-size tree: HKY85 + dGamma4, near-full-length reads over a 7 682-column
    consensus, windows of ~7 000 columns -- the warp-per-item estimate kernel (estimateWideKernel) and K4's global-memory
    rho slab -- against the oracle, whole pipeline from the reads."""
    from hmmufotu_b200 import engine
    db, otree, gdb = _setup(engine, n_leaves=160, cs_len=7682, occupancy=0.18, seed=20260105, model=synth.gg97_model(synth.HKY85), dg_alpha=0.5)
    rs = synth.make_reads(db, 12, read_len=1400, len_sd=50.0, seed=4, window=(0, 300))
    ohmm = O.OracleHmm(db)
    max_len = int(np.max(np.diff(rs.offsets))) + 8
    b = engine.Batch(gdb, rs.n, max_len)
    b.upload(rs)
    b.assign(engine.SepParams())
    rows, traces = b.alignments()
    ns, ids, dist = b.seeds(with_dist=True)
    nk, est = b.estimates()
    _, pl = b.places()
    best = b.best()
    widths = []
    for i in range(rs.n):
        o = ohmm.align(rs.read_str(i), rs.vpaths[i, :rs.n_paths[i]])
        assert rows[i]["cost"] == o["cost"] and traces[i] == o["trace"], i
        seq = O.digitize(o["align"])
        s, e = o["coords"][4] - 1, o["coords"][5] - 1
        widths.append(e - s + 1)
        _check_read(otree, seq, s, e, ns[i], ids[i], dist[i], nk[i], est[i], pl[i], best[i])
    assert min(widths) > 5000
    b.close(); gdb.close()


def test_paired_end_gtr_dgamma_merged_windows():
    """BASELINE.json configs[1] / [3]'s shape: GTR + dGamma4, 2 x 250 bp pairs with overlapping mates, merged windows wider
    than a single read's (the big-window configuration of the staged estimate kernel), whole pipeline against the oracle."""
    from hmmufotu_b200 import engine, _lib
    db, otree, gdb = _setup(engine, n_leaves=120, cs_len=3200, occupancy=0.3, seed=20260104, model=synth.gg97_model(synth.GTR), dg_alpha=0.5)
    fwd, rev = synth.make_pairs(db, 24, read_len=250, insert_len=420, insert_sd=20.0, seed=6)
    ohmm = O.OracleHmm(db)
    b = engine.Batch(gdb, fwd.n, 260, paired=True)
    b.upload(fwd, rev)
    b.assign(engine.SepParams())
    merged = b.merged()
    ns, ids, dist = b.seeds(with_dist=True)
    nk, est = b.estimates()
    _, pl = b.places()
    best = b.best()
    widths = []
    for i in range(fwd.n):
        a = ohmm.align(fwd.read_str(i), fwd.vpaths[i, :fwd.n_paths[i]])
        m = ohmm.align(rev.read_str(i), rev.vpaths[i, :rev.n_paths[i]])
        if not (a["coords"][4] <= m["coords"][4] and a["coords"][5] <= m["coords"][5]):
            assert merged[i]["flags"] & _lib.ALIGN_CHIMERA
            continue
        mm = O.merge(db.cs_len, a, m)
        assert merged[i]["cost"] == mm["cost"]
        seq = O.digitize(mm["align"])
        s, e = mm["coords"][4] - 1, mm["coords"][5] - 1
        widths.append(e - s + 1)
        _check_read(otree, seq, s, e, ns[i], ids[i], dist[i], nk[i], est[i], pl[i], best[i])
    assert len(widths) > fwd.n // 2 and max(widths) > 1200
    b.close(); gdb.close()
