"""CPU: the C-ABI library loads, exports every symbol include/hmmufotu_b200.h declares, its host-only
entry points agree with the oracle, and the compute entry points refuse to run without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_py as O
from hmmufotu_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from hmmufotu_b200 import build
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "hmmufotu_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(hu_[a-zA-Z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_lib.SYMBOLS)
    assert b"hmmufotu-b200" in lib.hu_version()


def test_struct_layouts_match_header():
    # sizes the kernels and the Python mirror must agree on (LP64)
    assert C.sizeof(_lib.AlignRow) == 48 and C.sizeof(_lib.PlaceRow) == 120
    assert C.sizeof(_lib.SepParams) == 56 and C.sizeof(_lib.ReadBatch) == 40


def test_hmm_finalize_matches_oracle(lib):
    from hmmufotu_b200 import engine
    db = synth.make_db(n_leaves=12, cs_len=300, occupancy=0.8, seed=2)
    entry, exit_ = engine.hmm_finalize(db.K, db.T)
    p = O.OracleHmm(db).params()
    assert np.array_equal(entry, p["entry"]) and np.array_equal(exit_, p["exit"])


@pytest.mark.parametrize("mtype", range(6))
def test_model_pr_host_matches_oracle(lib, mtype):
    from hmmufotu_b200 import engine
    m = synth.gg97_model(mtype)
    om = O.OracleModel(m)
    for v in (0.0, 1e-5, 0.02, 0.3, 1.7):
        np.testing.assert_allclose(engine.model_Pr(m, v), om.Pr(v), rtol=1e-12, atol=1e-15)


def test_trace_to_align_matches_oracle(lib):
    """hu_trace_to_align (host buildGlobalAlign) on the oracle's own traces, including inserts and the padding quirk."""
    db = synth.make_db(n_leaves=20, cs_len=500, occupancy=0.7, seed=8, leaf_ins=0.01)
    h = O.OracleHmm(db)
    rs = synth.make_reads(db, 60, read_len=160, seed=4, indel_prob=0.6)
    T = np.ascontiguousarray(db.T); eM = np.ascontiguousarray(db.eM); eI = np.ascontiguousarray(db.eI)
    p2cs = np.ascontiguousarray(db.p2cs, dtype=np.int32)
    entry = np.zeros(db.K + 1); exit_ = np.zeros(db.K + 1)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    hp = _lib.HmmParams(db.K, db.cs_len, vp(eM), vp(eI), vp(T), vp(entry), vp(exit_), vp(p2cs))
    n_ins = 0
    for i in range(rs.n):
        read = rs.read_str(i)
        a = h.align(read, rs.vpaths[i, :rs.n_paths[i]])
        row = _lib.AlignRow(*[int(x) for x in a["coords"]], 0, len(a["trace"]), a["cost"], 0)
        out = C.create_string_buffer(db.cs_len)
        assert lib.hu_trace_to_align(C.byref(hp), read.encode(), len(read), C.byref(row), a["trace"].encode(), out) == 0
        assert out.raw.decode() == a["align"], i
        n_ins += "I" in a["trace"]
    assert n_ins > 3


def test_no_cpu_fallback(lib):
    """Without a CUDA device hu_db_create must fail loudly (HU_ERR_CUDA), never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from hmmufotu_b200 import engine
    db = synth.make_db(n_leaves=6, cs_len=60, occupancy=0.9, seed=1)
    with pytest.raises(_lib.HuError) as e:
        engine.Database(db)
    assert "no CPU fallback" in str(e.value)


def test_seed_key_formula_is_exact():
    """hu_seed.cu makeKey(): floor(d * inv2N[N] + 2^-20) with inv2N = 2^31 / N correctly rounded is floor(d * 2^31 / N)."""
    rng = np.random.default_rng(3)
    Ns = list(range(1, 700)) + [int(x) for x in rng.integers(700, 65536, 400)]
    for N in Ns:
        inv2 = np.ldexp(np.float64(1.0) / np.float64(N), 31)
        m = N
        while m % 2 == 0:
            m //= 2
        ds = np.unique(np.concatenate([np.arange(0, N + 1) if N < 700 else rng.integers(0, N + 1, 64), np.arange(0, N + 1, m), [0, 1, N - 1, N]])).astype(np.int64)
        got = np.floor(ds.astype(np.float64) * inv2 + np.float64(2.0 ** -20)).astype(np.int64)
        want = (ds << 31) // N
        assert (got == want).all(), N


def _replay_desc(lib, keys):
    k = np.ascontiguousarray(keys, dtype=np.float64)
    perm = np.zeros(max(1, k.size), dtype=np.int32)
    assert lib.hu_debug_stdsort(k.ctypes.data_as(C.c_void_p), k.size, perm.ctypes.data_as(C.c_void_p)) == 0
    return perm[:k.size]


def test_stdsort_replay_is_libstdcxx_sort(lib):
    """hu_stdsort.cuh (what filterKernel / finalizeKernel and the heap-sort fallback of the literal seed order run on the
    device) against the REAL std::sort(rbegin, rend) of this toolchain (oracle): same permutation for every n = 1..64 with
    heavy ties, all-equal keys, NaN keys, sorted / reversed input, and -- well past 64 -- on key vectors built by McIlroy's
    adversary, on which introsort exhausts its 2 lg n depth budget and finishes with __partial_sort (heap sort)."""
    rng = np.random.default_rng(11)
    for n in range(1, 65):
        cases = [np.zeros(n), np.arange(n, dtype=float), np.arange(n, dtype=float)[::-1]]
        for vals in (2, 3, 5, n):
            cases += [rng.integers(0, vals, n).astype(float) for _ in range(6)]
        nan = rng.integers(0, 3, n).astype(float)
        nan[rng.random(n) < 0.3] = np.nan
        cases += [nan, np.full(n, np.nan)]
        for keys in cases:
            assert np.array_equal(_replay_desc(lib, keys), O.stdsort_desc_perm(keys)), (n, keys)
    assert [int(_replay_desc(lib, np.ones(n))[0]) for n in (17, 20, 33, 50)] == [15, 18, 31, 25]   # the equal-key artefact
    for n in (40, 64, 1000, 5000):
        killer = O.sort_killer(n).astype(float)[::-1].copy()    # reverse-iterator sort sees the mirrored array
        assert np.array_equal(_replay_desc(lib, killer), O.stdsort_desc_perm(killer)), n
        ties = rng.integers(0, 7, n).astype(float)
        assert np.array_equal(_replay_desc(lib, ties), O.stdsort_desc_perm(ties)), n
