"""ctypes driver for the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = _HERE
_LIB: Optional[C.CDLL] = None


class PlaceRow(C.Structure):
    _fields_ = [("start", C.c_int32), ("end", C.c_int32), ("c", C.c_int64), ("p", C.c_int64), ("a", C.c_int64),
                ("wuv", C.c_double), ("ratio", C.c_double), ("wnr", C.c_double), ("loglik", C.c_double),
                ("loglikTrue", C.c_double), ("height", C.c_double), ("qPlace", C.c_double), ("qTaxon", C.c_double),
                ("estLoglik", C.c_double), ("outerIters", C.c_int32), ("pad", C.c_int32)]


def build(force: bool = False) -> str:
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("oracle_hmm.cpp", "oracle_phylo.cpp", "oracle_capi.cpp", "oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs if os.path.exists(s)):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, i32, i64, f64 = C.c_void_p, C.c_int, C.c_long, C.c_double
        P = C.POINTER
        L.orc_hmm_from_costs.restype = vp
        L.orc_hmm_from_costs.argtypes = [i32, i32, vp, vp, vp, vp]
        L.orc_hmm_load.restype = vp
        L.orc_hmm_load.argtypes = [C.c_char_p]
        L.orc_hmm_free.argtypes = [vp]
        L.orc_hmm_set_mode.argtypes = [vp, i32]
        L.orc_hmm_K.argtypes = [vp]
        L.orc_hmm_L.argtypes = [vp]
        L.orc_hmm_get.argtypes = [vp] * 8
        L.orc_align.argtypes = [vp, C.c_char_p, i32, vp, i32, vp, P(f64), vp, vp, i32, P(i32), P(i64)]
        L.orc_merge.argtypes = [i32, vp, P(f64), vp, vp, f64, vp]
        L.orc_digitize.argtypes = [C.c_char_p, i32, vp]
        L.orc_model_create.restype = vp
        L.orc_model_create.argtypes = [i32, vp, vp, f64, f64, f64, f64]
        L.orc_model_load_text.restype = vp
        L.orc_model_load_text.argtypes = [C.c_char_p]
        L.orc_model_free.argtypes = [vp]
        L.orc_model_Pr.argtypes = [vp, f64, vp]
        L.orc_model_pi.argtypes = [vp, vp]
        L.orc_model_type.argtypes = [vp]
        L.orc_tree_create.restype = vp
        L.orc_tree_create.argtypes = [i64, i32, i64, vp, vp, vp, vp, vp, vp, i32, vp]
        L.orc_tree_free.argtypes = [vp]
        L.orc_tree_evaluate.argtypes = [vp, i32]
        L.orc_tree_get.argtypes = [vp, vp, vp, vp]
        L.orc_tree_set_messages.argtypes = [vp, vp, vp]
        L.orc_get_seed.argtypes = [vp, vp, i32, i32, f64, f64, i32, i32, vp, vp]
        L.orc_pdist.restype = i64
        L.orc_pdist.argtypes = [vp, i64, vp, i32, i32, P(i64)]
        L.orc_estimate.argtypes = [vp, vp, i32, i32, i64, f64, P(PlaceRow)]
        L.orc_place.argtypes = [vp, vp, P(PlaceRow), f64]
        L.orc_assign.argtypes = [vp, vp, i32, i32, i32, vp, vp, f64, f64, i32, P(PlaceRow)]
        L.orc_equal_sort_perm.argtypes = [i32, vp]
        L.orc_tree_set_options.argtypes = [vp, i32, i32, i32]
        L.orc_tree_evaluate_mt.argtypes = [vp, i32, i32]
        L.orc_run_reads.argtypes = [vp, vp, i32, i32, C.c_char_p, vp, vp, vp, i32, i32, f64, i32, i32, i32, P(PlaceRow), vp, vp, vp]
        L.orc_tree_set_sparse_messages.argtypes = [vp, i64, vp, vp, vp]
        L.orc_stdsort_desc_perm.argtypes = [vp, i32, vp]
        L.orc_stdsort_keys.argtypes = [vp, i32, i32, vp]
        L.orc_sort_killer.argtypes = [i32, vp]
        assert L.orc_sizeof_place_row() == C.sizeof(PlaceRow)
        _LIB = L
    return _LIB


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


GLOBAL, LOCAL, NGCL, CGNL = 0, 1, 2, 3


class OracleHmm:
    def __init__(self, db=None, path: Optional[str] = None, mode: int = GLOBAL):
        L = lib()
        if path is not None:
            self.h = L.orc_hmm_load(path.encode())
            if not self.h:
                raise IOError("oracle could not parse " + path)
        else:
            eM = np.ascontiguousarray(db.eM, dtype=np.float64)
            eI = np.ascontiguousarray(db.eI, dtype=np.float64)
            T = np.ascontiguousarray(db.T, dtype=np.float64)
            p2cs = np.ascontiguousarray(db.p2cs, dtype=np.int32)
            self.h = L.orc_hmm_from_costs(db.K, db.cs_len, _p(eM), _p(eI), _p(T), _p(p2cs))
        self.K = L.orc_hmm_K(self.h)
        self.L = L.orc_hmm_L(self.h)
        L.orc_hmm_set_mode(self.h, mode)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_hmm_free(self.h)
            self.h = None

    def params(self):
        K, Lc = self.K, self.L
        eM = np.zeros((K + 1, 4)); eI = np.zeros((K + 1, 4)); T = np.zeros((K + 1, 3, 3))
        entry = np.zeros(K + 1); exit_ = np.zeros(K + 1)
        p2cs = np.zeros(K + 1, dtype=np.int32); cs2p = np.zeros(Lc + 1, dtype=np.int32)
        lib().orc_hmm_get(self.h, _p(eM), _p(eI), _p(T), _p(entry), _p(exit_), _p(p2cs), _p(cs2p))
        return dict(eM=eM, eI=eI, T=T, entry=entry, exit=exit_, p2cs=p2cs, cs2p=cs2p)

    def align(self, read: str, vpaths: np.ndarray):
        """-> dict(coords[6], cost, align(str), trace(str), fallback, cells)"""
        vp = np.ascontiguousarray(vpaths, dtype=np.int32).reshape(-1, 6)
        coords = np.zeros(6, dtype=np.int32)
        cost = C.c_double()
        align = C.create_string_buffer(self.L)
        cap = len(read) + self.K + 8
        trace = C.create_string_buffer(cap)
        fb = C.c_int()
        cells = C.c_long()
        rc = lib().orc_align(self.h, read.encode(), len(read), _p(vp), vp.shape[0], _p(coords), C.byref(cost),
                             align, trace, cap, C.byref(fb), C.byref(cells))
        if rc < 0:
            raise RuntimeError("trace buffer too small")
        return dict(coords=coords, cost=cost.value, align=align.raw.decode(), trace=trace.value.decode(),
                    fallback=bool(fb.value), cells=cells.value, valid=rc == 0)


def merge(L: int, a: dict, b: dict) -> dict:
    coords = a["coords"].copy()
    cost = C.c_double(a["cost"])
    al = C.create_string_buffer(a["align"].encode(), L)
    oc = np.ascontiguousarray(b["coords"], dtype=np.int32)
    lib().orc_merge(L, _p(coords), C.byref(cost), al, _p(oc), b["cost"], b["align"].encode())
    return dict(coords=coords, cost=cost.value, align=al.raw.decode())


def digitize(align: str) -> np.ndarray:
    out = np.zeros(len(align), dtype=np.int8)
    n = lib().orc_digitize(align.encode(), len(align), _p(out))
    return out[:n]


class OracleModel:
    def __init__(self, m=None, text: Optional[str] = None):
        L = lib()
        if text is not None:
            self.h = L.orc_model_load_text(text.encode())
            if not self.h:
                raise ValueError("oracle could not parse the model text")
        else:
            pi = np.ascontiguousarray(m["pi"], dtype=np.float64)
            R = np.ascontiguousarray(m["R"], dtype=np.float64)
            self.h = L.orc_model_create(int(m["type"]), _p(pi), _p(R), float(m["kr"]), float(m["ky"]), float(m["kappa"]), float(m["beta"]))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_model_free(self.h)
            self.h = None

    def Pr(self, v: float) -> np.ndarray:
        P = np.zeros((4, 4))
        lib().orc_model_Pr(self.h, float(v), _p(P))
        return P

    def pi(self) -> np.ndarray:
        p = np.zeros(4)
        lib().orc_model_pi(self.h, _p(p))
        return p


class OracleTree:
    def __init__(self, db, evaluate: bool = True, infer_internal: bool = True, threads: int = 1):
        L = lib()
        self.db = db
        self.model = OracleModel(db.model)
        parent = np.ascontiguousarray(db.parent, dtype=np.int32)
        blen = np.ascontiguousarray(db.blen, dtype=np.float64)
        height = np.ascontiguousarray(db.height, dtype=np.float64)
        seq = np.ascontiguousarray(db.seq, dtype=np.int8)
        taxon = np.ascontiguousarray(db.taxon, dtype=np.int32)
        rates = np.ascontiguousarray(db.dg_rates, dtype=np.float64)
        self.n_nodes, self.cs_len = db.n_nodes, db.cs_len
        self.h = L.orc_tree_create(db.n_nodes, db.cs_len, db.root, _p(parent), _p(blen), _p(height), _p(seq), _p(taxon),
                                   self.model.h, int(rates.size), _p(rates) if rates.size else None)
        if evaluate:
            L.orc_tree_evaluate_mt(self.h, int(infer_internal), int(threads))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_tree_free(self.h)
            self.h = None

    def set_sparse_messages(self, ids: np.ndarray, up: np.ndarray, dn: np.ndarray):
        """Messages of the nodes `ids` only (row i of up / dn belongs to ids[i]); the other nodes have none."""
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        up = np.ascontiguousarray(up, dtype=np.float64); dn = np.ascontiguousarray(dn, dtype=np.float64)
        lib().orc_tree_set_sparse_messages(self.h, ids.size, _p(ids), _p(up), _p(dn))

    def set_options(self, weighted: bool = False, prior_height: bool = False, only_ml: bool = False):
        """-m weighted / --prior height / --ML of the reference's hmmufotu (src/hmmufotu.cpp:229-241)"""
        lib().orc_tree_set_options(self.h, int(weighted), int(prior_height), int(only_ml))

    def messages(self) -> Tuple[np.ndarray, np.ndarray]:
        up = np.zeros((self.n_nodes, self.cs_len, 4)); dn = np.zeros_like(up)
        lib().orc_tree_get(self.h, _p(up), _p(dn), None)
        return up, dn

    def seqs(self) -> np.ndarray:
        s = np.zeros((self.n_nodes, self.cs_len), dtype=np.int8)
        lib().orc_tree_get(self.h, None, None, _p(s))
        return s

    def get_seed(self, seq: np.ndarray, start: int, end: int, max_n: int = 50, max_diff: float = np.inf,
                 max_height: float = np.inf, order: int = 1):
        ids = np.zeros(max_n, dtype=np.int64); dist = np.zeros(max_n)
        s = np.ascontiguousarray(seq, dtype=np.int8)
        n = lib().orc_get_seed(self.h, _p(s), start, end, max_diff, max_height, order, max_n, _p(ids), _p(dist))
        return ids[:n], dist[:n]

    def pdist(self, node: int, seq: np.ndarray, start: int, end: int) -> Tuple[int, int]:
        N = C.c_long()
        s = np.ascontiguousarray(seq, dtype=np.int8)
        d = lib().orc_pdist(self.h, node, _p(s), start, end, C.byref(N))
        return int(d), int(N.value)

    def estimate(self, seq: np.ndarray, start: int, end: int, node: int, dist: float) -> PlaceRow:
        r = PlaceRow()
        s = np.ascontiguousarray(seq, dtype=np.int8)
        lib().orc_estimate(self.h, _p(s), start, end, node, dist, C.byref(r))
        return r

    def place(self, seq: np.ndarray, row: PlaceRow, max_height: float = np.inf) -> PlaceRow:
        s = np.ascontiguousarray(seq, dtype=np.int8)
        lib().orc_place(self.h, _p(s), C.byref(row), max_height)
        return row

    def assign(self, seq: np.ndarray, start: int, end: int, ids: Sequence[int], dist: Sequence[float],
               max_error: float = 20.0, max_height: float = np.inf, use_true: bool = False) -> List[PlaceRow]:
        ids = np.ascontiguousarray(ids, dtype=np.int64); dist = np.ascontiguousarray(dist, dtype=np.float64)
        rows = (PlaceRow * max(1, ids.size))()
        s = np.ascontiguousarray(seq, dtype=np.int8)
        n = lib().orc_assign(self.h, _p(s), start, end, ids.size, _p(ids), _p(dist), max_error, max_height, int(use_true), rows)
        return [rows[i] for i in range(n)]


def run_reads(hmm: "OracleHmm", tree: "OracleTree", reads, col0: int = 0, max_n_seed: int = 50, max_error: float = 20.0,
              seed_order: int = 0, use_true: bool = False, threads: int = 1, mates=None, with_ties: bool = False):
    """The reference's per-read task body for a ReadSet (and its mates, already reverse-complemented) on `threads` host
    threads.  -> (rows: list[PlaceRow], coords int32 [n, 6], cost f64 [n], n_bad[, near_tie int32 [n]])"""
    n = reads.n
    sets = (reads,) if mates is None else (reads, mates)
    text = "".join("".join("ACGT"[c] for c in x.codes) for x in sets).encode()
    offsets = np.ascontiguousarray(reads.offsets, dtype=np.int64)
    vp = np.ascontiguousarray(reads.vpaths, dtype=np.int32)
    npth = np.ascontiguousarray(reads.n_paths, dtype=np.int32)
    if mates is not None:
        offsets = np.concatenate([offsets, mates.offsets[1:].astype(np.int64) + offsets[-1]])
        vp = np.ascontiguousarray(np.concatenate([vp, mates.vpaths.astype(np.int32)]))
        npth = np.ascontiguousarray(np.concatenate([npth, mates.n_paths.astype(np.int32)]))
    rows = (PlaceRow * n)()
    coords = np.zeros((n, 6), dtype=np.int32)
    cost = np.zeros(n)
    ties = np.zeros(n, dtype=np.int32)
    bad = lib().orc_run_reads(hmm.h, tree.h, n, int(mates is not None), text, _p(offsets), _p(vp), _p(npth), col0, max_n_seed, max_error,
                              seed_order, int(use_true), threads, rows, _p(coords), _p(cost), _p(ties))
    return (rows, coords, cost, bad, ties) if with_ties else (rows, coords, cost, bad)


def equal_sort_perm(n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.int32)
    lib().orc_equal_sort_perm(n, _p(out))
    return out


def stdsort_desc_perm(keys: np.ndarray) -> np.ndarray:
    """The real std::sort(rbegin, rend, byKey): index of the element left at each position."""
    k = np.ascontiguousarray(keys, dtype=np.float64)
    out = np.zeros(k.size, dtype=np.int32)
    lib().orc_stdsort_desc_perm(_p(k), k.size, _p(out))
    return out


def stdsort_keys(keys: np.ndarray, want: int) -> np.ndarray:
    """The real std::sort(begin, end) of getSeed on integer keys: candidate positions of the first `want` elements."""
    k = np.ascontiguousarray(keys, dtype=np.uint32)
    out = np.zeros(min(want, k.size), dtype=np.int32)
    lib().orc_stdsort_keys(_p(k), k.size, want, _p(out))
    return out


def sort_killer(n: int) -> np.ndarray:
    """Key vector on which this libstdc++'s std::sort exhausts introsort's depth budget (McIlroy's adversary)."""
    out = np.zeros(n, dtype=np.uint32)
    lib().orc_sort_killer(n, _p(out))
    return out
