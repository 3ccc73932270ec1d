"""Host-side mirror of the reference's per-read interface over the C ABI.

`Database` stands for the (hmm, ptu) pair `hmmufotu` loads (/root/reference/src/hmmufotu.cpp:457-494);
`Batch` exposes the stages of the per-read pipeline under the reference's names
(/root/reference/src/HmmUFOtu_main.h:70-113): alignSeq, getSeed, estimateSeq (+filterPlacements), placeSeq,
calcQValues, batched.  All arithmetic happens in libhmmufotu_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np

from . import _lib
from ._lib import check

ALIGN_DT = np.dtype([("seqStart", "<i4"), ("seqEnd", "<i4"), ("hmmStart", "<i4"), ("hmmEnd", "<i4"), ("csStart", "<i4"),
                     ("csEnd", "<i4"), ("flags", "<i4"), ("traceLen", "<i4"), ("cost", "<f8"), ("cells", "<i8")])
PLACE_DT = np.dtype([("start", "<i4"), ("end", "<i4"), ("cNode", "<i4"), ("pNode", "<i4"), ("aNode", "<i4"), ("taxon", "<i4"),
                     ("dist", "<f8"), ("wuv", "<f8"), ("ratio", "<f8"), ("wnr", "<f8"), ("loglik", "<f8"),
                     ("loglikTrue", "<f8"), ("estLoglik", "<f8"), ("height", "<f8"), ("annoDist", "<f8"),
                     ("qPlace", "<f8"), ("qTaxon", "<f8"), ("outerIters", "<i4"), ("nPlaces", "<i4")])
assert ALIGN_DT.itemsize == C.sizeof(_lib.AlignRow) and PLACE_DT.itemsize == C.sizeof(_lib.PlaceRow)


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def model_desc(m: Dict, dg_rates=None) -> _lib.ModelDesc:
    d = _lib.ModelDesc()
    d.type = int(m["type"])
    rates = np.zeros(0) if dg_rates is None else np.asarray(dg_rates, dtype=np.float64)
    d.dgK = int(rates.size)
    for i in range(4):
        d.pi[i] = float(m["pi"][i])
    R = np.asarray(m.get("R", np.zeros((4, 4))), dtype=np.float64).reshape(16)
    for i in range(16):
        d.R[i] = float(R[i])
    d.kr, d.ky, d.kappa, d.beta = float(m["kr"]), float(m["ky"]), float(m["kappa"]), float(m["beta"])
    for i in range(rates.size):
        d.dgRates[i] = float(rates[i])
    return d


def model_Pr(m: Dict, v: float) -> np.ndarray:
    """DNASubModel::Pr(v) through the library (host evaluation of the kernels' code)."""
    P = np.zeros((4, 4))
    d = model_desc(m)
    check(_lib.load().hu_model_Pr(C.byref(d), float(v), _p(P)))
    return P


def hmm_finalize(K: int, T: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    T = np.ascontiguousarray(T, dtype=np.float64)
    entry = np.zeros(K + 1)
    exit_ = np.zeros(K + 1)
    check(_lib.load().hu_hmm_finalize(K, _p(T), _p(entry), _p(exit_)))
    return entry, exit_


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array over page-locked host memory (cudaMallocHost); freed when the array is collected."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    L = _lib.load()
    ptr = L.hu_host_alloc(max(n, 1))
    if not ptr:
        raise _lib.HuError("hu_host_alloc failed: " + L.hu_last_error().decode())
    buf = (C.c_char * max(n, 1)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[ptr] = buf
    import weakref
    weakref.finalize(buf, L.hu_host_free, ptr)
    return arr


_PINNED = {}


def pinned_copy(a: np.ndarray) -> np.ndarray:
    out = pinned_empty(a.shape, a.dtype)
    out[...] = a
    return out


class Database:
    """One database replica on one GPU."""

    def __init__(self, db, device: int = 0, messages: Optional[Tuple[np.ndarray, np.ndarray]] = None,
                 infer_internal: bool = True):
        L = _lib.load()
        self.src = db
        self.K, self.L = int(db.K), int(db.cs_len)
        self.n_nodes = int(db.n_nodes)
        self._keep = []
        eM = np.ascontiguousarray(db.eM, dtype=np.float64)
        eI = np.ascontiguousarray(db.eI, dtype=np.float64)
        T = np.ascontiguousarray(db.T, dtype=np.float64)
        p2cs = np.ascontiguousarray(db.p2cs, dtype=np.int32)
        entry, exit_ = hmm_finalize(self.K, T)
        self.entry, self.exit = entry, exit_
        hp = _lib.HmmParams(self.K, self.L, _p(eM), _p(eI), _p(T), _p(entry), _p(exit_), _p(p2cs))
        self._hmm_arrays = (eM, eI, T, entry, exit_, p2cs)
        self.hmm_params = hp
        parent = np.ascontiguousarray(db.parent, dtype=np.int32)
        blen = np.ascontiguousarray(db.blen, dtype=np.float64)
        height = np.ascontiguousarray(db.height, dtype=np.float64)
        seq = np.ascontiguousarray(db.seq, dtype=np.int8)
        taxon = np.ascontiguousarray(db.taxon, dtype=np.int32)
        td = _lib.TreeDesc()
        td.nNodes, td.csLen, td.root = self.n_nodes, self.L, int(db.root)
        td.parent, td.blen, td.height, td.seq, td.taxon = _p(parent), _p(blen), _p(height), _p(seq), _p(taxon)
        if messages is not None:
            up = np.ascontiguousarray(messages[0], dtype=np.float64)
            dn = np.ascontiguousarray(messages[1], dtype=np.float64)
            td.up, td.dn = _p(up), _p(dn)
            td.inferInternalSeq = 0
        else:
            td.up = td.dn = None
            td.inferInternalSeq = int(infer_internal)
        md = model_desc(db.model, db.dg_rates)
        h = C.c_void_p()
        check(L.hu_db_create(C.byref(h), C.byref(hp), C.byref(td), C.byref(md), device))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            _lib.load().hu_db_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    @property
    def bytes(self) -> int:
        return int(_lib.load().hu_db_bytes(self.h))

    def messages(self, node0: int = 0, n: Optional[int] = None):
        n = self.n_nodes - node0 if n is None else n
        up = np.zeros((n, self.L, 4))
        dn = np.zeros((n, self.L, 4))
        check(_lib.load().hu_db_get_messages(self.h, node0, n, _p(up), _p(dn)))
        return up, dn

    def messages_of(self, ids, col0: int = 0, n_cols: Optional[int] = None):
        """(up, dn) of the listed nodes over columns [col0, col0 + n_cols): [len(ids), n_cols, 4], log domain."""
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        n_cols = self.L - col0 if n_cols is None else n_cols
        up = np.zeros((ids.size, n_cols, 4))
        dn = np.zeros((ids.size, n_cols, 4))
        check(_lib.load().hu_db_get_messages_of(self.h, ids.size, _p(ids), col0, n_cols, _p(up), _p(dn)))
        return up, dn

    def seqs(self, node0: int = 0, n: Optional[int] = None) -> np.ndarray:
        n = self.n_nodes - node0 if n is None else n
        s = np.zeros((n, self.L), dtype=np.int8)
        check(_lib.load().hu_db_get_seq(self.h, node0, n, _p(s)))
        return s

    def trace_to_align(self, read: str, row, trace: str) -> str:
        """BandedHMMP7::buildGlobalAlign on the host."""
        ar = _lib.AlignRow()
        for f, _ in _lib.AlignRow._fields_:
            setattr(ar, f, row[f].item() if hasattr(row[f], "item") else row[f])
        out = C.create_string_buffer(self.L)
        check(_lib.load().hu_trace_to_align(C.byref(self.hmm_params), read.encode(), len(read), C.byref(ar), trace.encode(), out))
        return out.raw.decode()


class SepParams:
    def __init__(self, max_n_seed: int = 50, max_error: float = 20.0, max_diff: float = np.inf,
                 max_height: float = np.inf, use_true_loglik: bool = False, seed_order: int = _lib.SEED_ORDER_STDSORT,
                 weighted: bool = False, prior_height: bool = False, only_ml: bool = False):
        p = _lib.SepParams()
        _lib.load().hu_sep_params_default(C.byref(p))
        p.maxNSeed, p.maxError, p.maxDiff, p.maxHeight = max_n_seed, max_error, max_diff, max_height
        p.useTrueLoglik = int(use_true_loglik)
        p.seedOrder = int(seed_order)
        p.estMethod, p.prior, p.onlyML = int(weighted), int(prior_height), int(only_ml)
        self.c = p


class Batch:
    """Device workspace for one batch of reads; stages named after the reference's functions."""

    def __init__(self, database: Database, max_reads: int, max_len: int, paired: bool = False):
        self.db = database
        self.max_reads, self.max_len, self.paired = max_reads, max_len, paired
        h = C.c_void_p()
        check(_lib.load().hu_batch_create(C.byref(h), database.h, max_reads, max_len, int(paired)))
        self.h = h
        self.n = 0
        self.n_seq = 0
        self.max_n_seed = 50
        self._keep = None
        self._res = None       # page-locked result buffers of alignments()
        self._best = None

    def close(self):
        if getattr(self, "h", None):
            _lib.load().hu_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    # ---- staging ----
    def upload(self, reads, mates=None):
        """reads (and mates, already reverse-complemented): synth.ReadSet-like with codes/offsets/vpaths/n_paths."""
        if (mates is not None) != self.paired:
            raise ValueError("pairedness differs from the workspace")
        if mates is None:
            codes, offsets, vp, npth = reads.codes, reads.offsets, reads.vpaths, reads.n_paths
        else:
            codes = np.concatenate([reads.codes, mates.codes])
            offsets = np.concatenate([reads.offsets, mates.offsets[1:] + reads.offsets[-1]])
            vp = np.concatenate([reads.vpaths, mates.vpaths])
            npth = np.concatenate([reads.n_paths, mates.n_paths])
        codes = np.ascontiguousarray(codes, dtype=np.int8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        vp = np.ascontiguousarray(vp, dtype=np.int32)
        npth = np.ascontiguousarray(npth, dtype=np.int32)
        self._keep = (codes, offsets, vp, npth)
        rb = _lib.ReadBatch(reads.n, int(self.paired), _p(codes), _p(offsets), _p(vp), _p(npth))
        check(_lib.load().hu_batch_upload(self.h, C.byref(rb)))
        self.n = reads.n
        self.n_seq = reads.n * (2 if self.paired else 1)

    def set_aligned(self, codes: np.ndarray, start: np.ndarray, end: np.ndarray):
        codes = np.ascontiguousarray(codes, dtype=np.int8)
        start = np.ascontiguousarray(start, dtype=np.int32)
        end = np.ascontiguousarray(end, dtype=np.int32)
        self._keep = (codes, start, end)
        check(_lib.load().hu_batch_set_aligned(self.h, codes.shape[0], _p(codes), _p(start), _p(end)))
        self.n = self.n_seq = codes.shape[0]

    # ---- stages ----
    def alignSeq(self, mode: int = _lib.GLOBAL):
        check(_lib.load().hu_align_batch(self.h, mode))

    def getSeed(self, p: SepParams):
        self.max_n_seed = p.c.maxNSeed
        check(_lib.load().hu_seed_batch(self.h, C.byref(p.c)))

    def estimateSeq(self, p: SepParams):
        check(_lib.load().hu_estimate_batch(self.h, C.byref(p.c)))

    def placeSeq(self, p: SepParams):
        check(_lib.load().hu_place_batch(self.h, C.byref(p.c)))

    def calcQValues(self, p: SepParams):
        check(_lib.load().hu_finalize_batch(self.h, C.byref(p.c)))

    def assign(self, p: SepParams, mode: int = _lib.GLOBAL):
        self.max_n_seed = p.c.maxNSeed
        check(_lib.load().hu_assign_batch(self.h, mode, C.byref(p.c)))

    def sep(self, p: SepParams):
        self.getSeed(p)
        self.estimateSeq(p)
        self.placeSeq(p)
        self.calcQValues(p)

    def sync(self):
        check(_lib.load().hu_batch_sync(self.h))

    # ---- results ----
    def alignments(self, with_traces=True):
        """-> (rows, traces).  with_traces True: list of "B[MID]*E" strings; "packed": (uint8 buffer, int64 offsets) as the
        library hands them back (NUL-terminated, back to back), in page-locked buffers owned by this workspace and reused by
        the next call; False: None."""
        if with_traces is False or with_traces is None:
            rows = np.zeros(self.n_seq, dtype=ALIGN_DT)
            check(_lib.load().hu_batch_get_align(self.h, _p(rows), None, 0, None))
            return rows, None
        if self._res is None:
            max_seq = self.max_reads * (2 if self.paired else 1)
            cap = int(max_seq) * (self.max_len + self.db.K + 8)
            self._res = (pinned_empty(max_seq, ALIGN_DT), pinned_empty(cap, np.uint8), pinned_empty(max_seq + 1, np.int64))
        rows_p, buf, off = self._res
        check(_lib.load().hu_batch_get_align(self.h, _p(rows_p), _p(buf), buf.size, _p(off)))
        rows = rows_p[:self.n_seq]
        if with_traces == "packed":
            return rows, (buf, off[:self.n_seq + 1])
        rows = rows.copy()
        raw = buf[:int(off[self.n_seq])].tobytes()
        traces = [raw[off[i]:off[i] + rows["traceLen"][i]].decode() for i in range(self.n_seq)]
        return rows, traces

    def merged(self) -> np.ndarray:
        rows = np.zeros(self.n, dtype=ALIGN_DT)
        check(_lib.load().hu_batch_get_merged(self.h, _p(rows)))
        return rows

    def aligned_codes(self) -> np.ndarray:
        codes = np.zeros((self.n, self.db.L), dtype=np.int8)
        check(_lib.load().hu_batch_get_aligned_codes(self.h, _p(codes)))
        return codes

    def seeds(self, with_dist: bool = False):
        n = np.zeros(self.n, dtype=np.int32)
        ids = np.zeros((self.n, self.max_n_seed), dtype=np.int32)
        dist = np.zeros((self.n, self.max_n_seed)) if with_dist else None
        check(_lib.load().hu_batch_get_seeds(self.h, _p(n), _p(ids), _p(dist)))
        return n, ids, dist

    def seed_keys(self, read: int) -> np.ndarray:
        """Diagnostic: the integer p-distance keys of one read over getSeed's candidate vector (after getSeed)."""
        n = C.c_int32()
        check(_lib.load().hu_batch_get_seed_keys(self.h, read, None, C.byref(n)))
        keys = np.zeros(n.value, dtype=np.uint32)
        check(_lib.load().hu_batch_get_seed_keys(self.h, read, _p(keys), C.byref(n)))
        return keys

    def estimates(self):
        n = np.zeros(self.n, dtype=np.int32)
        rows = np.zeros((self.n, self.max_n_seed), dtype=PLACE_DT)
        check(_lib.load().hu_batch_get_estimates(self.h, _p(n), _p(rows)))
        return n, rows

    def places(self):
        n = np.zeros(self.n, dtype=np.int32)
        rows = np.zeros((self.n, self.max_n_seed), dtype=PLACE_DT)
        check(_lib.load().hu_batch_get_places(self.h, _p(n), _p(rows)))
        return n, rows

    def best(self, pinned: bool = False) -> np.ndarray:
        """bestPlace of every read.  pinned: into a page-locked buffer of this workspace (reused by the next call)."""
        if pinned:
            if self._best is None:
                self._best = pinned_empty(self.max_reads, PLACE_DT)
            rows = self._best[:self.n]
        else:
            rows = np.zeros(self.n, dtype=PLACE_DT)
        check(_lib.load().hu_batch_get_best(self.h, _p(rows)))
        return rows

    def pack_rows(self, read0: int, dev_ptr: int):
        """The driver's fixed-size assignment rows of this batch, written on the batch stream into device memory."""
        check(_lib.load().hu_batch_pack_rows(self.h, int(read0), C.c_void_p(dev_ptr)))

    @property
    def stream(self) -> int:
        return int(_lib.load().hu_batch_stream(self.h) or 0)

    def mark(self, i: int):
        check(_lib.load().hu_batch_mark(self.h, i))

    def elapsed_ms(self, i: int, other: "Batch", j: int) -> float:
        """ms from mark i of this workspace to mark j of `other` (waits for the latter)."""
        ms = C.c_float()
        check(_lib.load().hu_batch_elapsed(self.h, i, other.h, j, C.byref(ms)))
        return float(ms.value)

    def stage_ms(self) -> np.ndarray:
        ms = np.zeros(8, dtype=np.float32)
        check(_lib.load().hu_batch_stage_ms(self.h, _p(ms)))
        return ms

    def stats(self) -> Dict[str, int]:
        s = _lib.BatchStats()
        check(_lib.load().hu_batch_get_stats(self.h, C.byref(s)))
        return {f: int(getattr(s, f)) for f, _ in _lib.BatchStats._fields_}
