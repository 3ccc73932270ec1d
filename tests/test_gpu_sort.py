"""-m gpu: the three std::sort calls of the per-read path (seeds: src/HmmUFOtu_main.cpp:139; estimated placements: :164;
final placements: src/hmmufotu.cpp:725-731) as the device replays them, against the REAL std::sort of this toolchain's
libstdc++ (oracle: orc_stdsort_keys / orc_stdsort_desc_perm) on the same keys -- at the benchmark's candidate count
(198 642 keys, i.e. the global-memory partition levels of seedSortLiteralKernel), on adversarial inputs that drive introsort
into its heap-sort fallback, and on exactly tied log-likelihoods (duplicated sibling leaves)."""
import ctypes as C

import numpy as np
import pytest

import oracle_py as O
from hmmufotu_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def _literal(keys: np.ndarray, want: int):
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    rows, n = keys.shape
    out = np.zeros((rows, want), dtype=np.int32)
    n_out = np.zeros(rows, dtype=np.int32)
    heap = np.zeros(1, dtype=np.int32)
    _lib.check(_lib.load().hu_debug_literal_order(0, keys.ctypes.data_as(C.c_void_p), rows, n, want, out.ctypes.data_as(C.c_void_p),
                                                  n_out.ctypes.data_as(C.c_void_p), heap.ctypes.data_as(C.c_void_p)))
    return out, n_out, int(heap[0])


def _pdist_like_keys(rng, n, n_cols=230, close=300):
    """Keys shaped like a read's p-distances to ~200 k nodes: small rationals d/N, most nodes far, a few hundred close,
    heavy ties everywhere (N takes a handful of values)."""
    N = rng.integers(n_cols - 6, n_cols + 1, n)
    p = np.where(rng.random(n) < close / n, rng.uniform(0.0, 0.05, n), rng.uniform(0.12, 0.35, n))
    d = rng.binomial(N, p)
    keys = np.floor(d * 2.0 ** 31 / N).astype(np.uint32)
    keys[rng.random(n) < 1e-4] = 0xFFFFFFFF            # undefined distances (0/0)
    return keys


def test_literal_seed_order_at_benchmark_scale():
    """198 642 candidates per read (BASELINE.json configs[2]): the first 50 positions std::sort leaves, 48 reads."""
    rng = np.random.default_rng(2026)
    n, rows, want = 198642, 48, 50
    keys = np.stack([_pdist_like_keys(rng, n, n_cols=int(rng.integers(180, 251)), close=int(rng.integers(30, 2000))) for _ in range(rows)])
    keys[0] = 12345                                     # all equal
    keys[1] = np.arange(n, dtype=np.uint32)             # sorted
    keys[2] = np.arange(n, dtype=np.uint32)[::-1]       # reversed
    keys[3] = rng.integers(0, 3, n)                     # three values
    got, n_out, heap = _literal(keys, want)
    for r in range(rows):
        ref = O.stdsort_keys(keys[r], want)
        k = int(n_out[r])
        assert k == want - int((keys[r][ref] == 0xFFFFFFFF).sum()), r
        assert np.array_equal(got[r, :k], ref[:k]), r
    assert heap == 0


@pytest.mark.parametrize("n", [1, 2, 3, 16, 17, 33, 100, 257, 1024, 1025, 4097, 8192, 8193, 20000])
def test_literal_seed_order_sizes(n):
    rng = np.random.default_rng(n)
    rows = 12
    keys = np.stack([rng.integers(0, max(2, n // (1 + r)), n).astype(np.uint32) for r in range(rows)])
    for want in (1, 50, 64):
        got, n_out, _ = _literal(keys, want)
        for r in range(rows):
            ref = O.stdsort_keys(keys[r], want)
            assert int(n_out[r]) == len(ref) and np.array_equal(got[r, :len(ref)], ref), (n, want, r)


@pytest.mark.parametrize("n", [1000, 5000, 20000])
def test_literal_seed_order_heap_sort_fallback(n):
    """Key vectors built by McIlroy's adversary against this libstdc++: the quicksort degenerates, __introsort_loop spends
    its 2 lg n depth budget inside the first 64 positions and finishes the range with __partial_sort (heap sort).  The
    device replays that too (n = 1000, 5000: in shared memory; 20000: in global memory)."""
    killer = O.sort_killer(n)
    keys = np.stack([killer, killer[::-1].copy(), np.roll(killer, 7)])
    got, n_out, heap = _literal(keys, 64)
    assert heap >= 1
    for r in range(keys.shape[0]):
        ref = O.stdsort_keys(keys[r], 64)
        assert int(n_out[r]) == 64 and np.array_equal(got[r], ref), (n, r)


@pytest.fixture(scope="module")
def tied():
    """A small database in which several sibling leaf pairs are exact duplicates (same sequence, same branch length): the
    two edges of such a pair give bit-identical estimated and final log-likelihoods."""
    from hmmufotu_b200 import engine
    db = synth.make_db(n_leaves=70, cs_len=600, occupancy=0.9, seed=23)
    ch = {}
    for x in range(db.n_nodes):
        if db.parent[x] >= 0:
            ch.setdefault(int(db.parent[x]), []).append(x)
    pairs = [(c[0], c[1]) for c in ch.values() if len(c) == 2 and db.is_leaf[c[0]] and db.is_leaf[c[1]]]
    assert len(pairs) >= 8
    for a, b in pairs:
        db.seq[b] = db.seq[a]
        db.blen[b] = db.blen[a]
        db.height[b] = db.height[a]
    otree = O.OracleTree(db)
    up, dn = otree.messages()
    db.seq = otree.seqs()
    gdb = engine.Database(db, messages=(up, dn))
    reads = synth.make_reads(db, 256, read_len=180, seed=9)
    return dict(db=db, otree=otree, gdb=gdb, reads=reads, engine=engine, pairs=pairs)


@pytest.mark.parametrize("variant", ["default", "ml_true", "height", "max_height"])
def test_placement_sorts_are_std_sort_on_exact_ties(tied, variant):
    """filterPlacements' order and the reported placement = what the real std::sort(rbegin, rend) does with the device's own
    keys (n = 50 > 16: introsort, unstable), including exact ties; under -H the literal seed order runs on the
    height-filtered candidate vector (src/HmmUFOtu_main.cpp:131-139) and equals the oracle's getSeed."""
    engine, rs, gdb, otree, db = tied["engine"], tied["reads"], tied["gdb"], tied["otree"], tied["db"]
    kw = dict(default={}, ml_true=dict(only_ml=True, use_true_loglik=True), height=dict(prior_height=True, use_true_loglik=True),
              max_height=dict(max_height=float(np.quantile(db.height, 0.8))))[variant]
    p = engine.SepParams(max_error=1e9, **kw)            # keep every estimate: the whole permutation is observable
    b = engine.Batch(gdb, rs.n, 200)
    b.upload(rs)
    b.assign(p)
    ns, ids, dist = b.seeds(with_dist=True)
    nk, est = b.estimates()
    _, pl = b.places()
    best = b.best()
    codes = b.aligned_codes()
    merged = b.merged()
    n_tie_reads = 0
    for i in range(rs.n):
        if variant == "max_height":
            s, e = merged["csStart"][i] - 1, merged["csEnd"][i] - 1
            oid, odist = otree.get_seed(codes[i], s, e, 50, max_height=kw["max_height"], order=0)
            assert ns[i] == len(oid) and np.array_equal(ids[i, :ns[i]], oid), i
            assert (db.height[ids[i, :ns[i]]] <= kw["max_height"]).all()
        # estimates come back in filter order; their keys in SEED order are what std::sort saw
        n = int(ns[i])
        assert nk[i] == n
        kept = est[i, :n]
        key_of = {int(r["cNode"]): float(r["estLoglik"]) for r in kept}
        keys = np.array([key_of[int(c)] for c in ids[i, :n]])
        perm = O.stdsort_desc_perm(keys)
        assert [int(c) for c in kept["cNode"]] == [int(ids[i, k]) for k in perm], i
        n_tie_reads += len(set(keys.tolist())) < n
        # final ordering: keys of the placements in filter order
        q = pl[i, :nk[i]]
        fkeys = (q["loglik"] if variant == "ml_true" else q["qPlace"]).astype(float)
        first = int(O.stdsort_desc_perm(fkeys)[0])
        assert (best[i]["cNode"], best[i]["pNode"]) == (q[first]["cNode"], q[first]["pNode"]), (variant, i)
    assert n_tie_reads > 0                            # exact ties were really present in fully observable orders
    b.close()
