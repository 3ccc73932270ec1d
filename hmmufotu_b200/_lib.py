"""ctypes binding of libhmmufotu_b200.so (include/hmmufotu_b200.h).  No fallback: a missing library or a
missing CUDA device is an error."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libhmmufotu_b200.so")

HU_MAX_SEED = 64
HU_MAX_DG = 8
GLOBAL, LOCAL, NGCL, CGNL = 0, 1, 2, 3
SEED_ORDER_STDSORT, SEED_ORDER_ID = 0, 1
ALIGN_FALLBACK, ALIGN_NOSEED, ALIGN_IRREGULAR, ALIGN_CHIMERA, ALIGN_INVALID = 1, 2, 4, 8, 16


class HmmParams(C.Structure):
    _fields_ = [("K", C.c_int32), ("L", C.c_int32), ("eM", C.c_void_p), ("eI", C.c_void_p), ("T", C.c_void_p),
                ("entry", C.c_void_p), ("exit", C.c_void_p), ("p2cs", C.c_void_p)]


class ModelDesc(C.Structure):
    _fields_ = [("type", C.c_int32), ("dgK", C.c_int32), ("pi", C.c_double * 4), ("R", C.c_double * 16),
                ("kr", C.c_double), ("ky", C.c_double), ("kappa", C.c_double), ("beta", C.c_double),
                ("dgRates", C.c_double * HU_MAX_DG)]


class TreeDesc(C.Structure):
    _fields_ = [("nNodes", C.c_int64), ("csLen", C.c_int32), ("root", C.c_int32), ("parent", C.c_void_p),
                ("blen", C.c_void_p), ("height", C.c_void_p), ("seq", C.c_void_p), ("taxon", C.c_void_p),
                ("up", C.c_void_p), ("dn", C.c_void_p), ("inferInternalSeq", C.c_int32)]


class ReadBatch(C.Structure):
    _fields_ = [("nReads", C.c_int32), ("paired", C.c_int32), ("codes", C.c_void_p), ("offsets", C.c_void_p),
                ("vpaths", C.c_void_p), ("nPaths", C.c_void_p)]


class AlignRow(C.Structure):
    _fields_ = [("seqStart", C.c_int32), ("seqEnd", C.c_int32), ("hmmStart", C.c_int32), ("hmmEnd", C.c_int32),
                ("csStart", C.c_int32), ("csEnd", C.c_int32), ("flags", C.c_int32), ("traceLen", C.c_int32),
                ("cost", C.c_double), ("cells", C.c_int64)]


class PlaceRow(C.Structure):
    _fields_ = [("start", C.c_int32), ("end", C.c_int32), ("cNode", C.c_int32), ("pNode", C.c_int32),
                ("aNode", C.c_int32), ("taxon", C.c_int32), ("dist", C.c_double), ("wuv", C.c_double),
                ("ratio", C.c_double), ("wnr", C.c_double), ("loglik", C.c_double), ("loglikTrue", C.c_double),
                ("estLoglik", C.c_double), ("height", C.c_double), ("annoDist", C.c_double), ("qPlace", C.c_double),
                ("qTaxon", C.c_double), ("outerIters", C.c_int32), ("nPlaces", C.c_int32)]


class SepParams(C.Structure):
    _fields_ = [("maxNSeed", C.c_int32), ("useTrueLoglik", C.c_int32), ("maxDiff", C.c_double),
                ("maxHeight", C.c_double), ("maxError", C.c_double), ("equalSortPerm", C.c_void_p),
                ("seedOrder", C.c_int32), ("estMethod", C.c_int32), ("prior", C.c_int32), ("onlyML", C.c_int32)]


class BatchStats(C.Structure):
    _fields_ = [("cells", C.c_int64), ("seedPairs", C.c_int64), ("seedBytes", C.c_int64), ("estItems", C.c_int64),
                ("estBytes", C.c_int64), ("placeItems", C.c_int64), ("placeBytes", C.c_int64), ("launches", C.c_int64),
                ("seedHeapSorts", C.c_int64), ("seedMacs", C.c_int64), ("seedScanFlags", C.c_int64)]


# every symbol include/hmmufotu_b200.h declares
SYMBOLS = ["hu_hmm_finalize", "hu_model_Pr", "hu_db_create", "hu_db_destroy", "hu_db_get_messages", "hu_db_get_seq",
           "hu_db_bytes", "hu_sep_params_default", "hu_batch_create", "hu_batch_destroy", "hu_batch_upload",
           "hu_batch_set_aligned", "hu_align_batch", "hu_seed_batch", "hu_estimate_batch", "hu_place_batch",
           "hu_finalize_batch", "hu_assign_batch", "hu_batch_sync", "hu_batch_get_align", "hu_batch_get_merged",
           "hu_batch_get_aligned_codes", "hu_batch_get_seeds", "hu_batch_get_estimates", "hu_batch_get_places",
           "hu_batch_get_best", "hu_batch_stage_ms", "hu_batch_get_stats", "hu_trace_to_align", "hu_last_error",
           "hu_version", "hu_host_alloc", "hu_host_free", "hu_batch_mark", "hu_batch_elapsed", "hu_batch_get_seed_keys",
           "hu_debug_literal_order", "hu_debug_stdsort", "hu_batch_pack_rows", "hu_batch_stream", "hu_db_get_messages_of", "hu_filter_batch", "hu_batch_set_seeds",
           "hu_batch_set_estimates", "hu_batch_set_places", "hu_db_create_deferred", "hu_db_put_messages", "hu_db_finish", "hu_db_clone"]

_LIB = None


def load() -> C.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s is missing: run `python -m hmmufotu_b200.build` (there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    P = C.POINTER
    L.hu_last_error.restype = C.c_char_p
    L.hu_version.restype = C.c_char_p
    L.hu_hmm_finalize.argtypes = [i32, vp, vp, vp]
    L.hu_model_Pr.argtypes = [P(ModelDesc), f64, vp]
    L.hu_db_create.argtypes = [P(vp), P(HmmParams), P(TreeDesc), P(ModelDesc), C.c_int]
    L.hu_db_destroy.argtypes = [vp]
    L.hu_db_create_deferred.argtypes = [P(vp), P(HmmParams), P(TreeDesc), P(ModelDesc), C.c_int]
    L.hu_db_put_messages.argtypes = [vp, i64, vp, vp, vp]
    L.hu_db_finish.argtypes = [vp]
    L.hu_db_clone.argtypes = [P(vp), vp, C.c_int]
    L.hu_db_get_messages.argtypes = [vp, i64, i64, vp, vp]
    L.hu_db_get_seq.argtypes = [vp, i64, i64, vp]
    L.hu_db_get_messages_of.argtypes = [vp, i64, vp, i32, i32, vp, vp]
    L.hu_db_bytes.restype = i64
    L.hu_db_bytes.argtypes = [vp]
    L.hu_sep_params_default.argtypes = [P(SepParams)]
    L.hu_batch_create.argtypes = [P(vp), vp, i32, i32, i32]
    L.hu_batch_destroy.argtypes = [vp]
    L.hu_batch_upload.argtypes = [vp, P(ReadBatch)]
    L.hu_batch_set_aligned.argtypes = [vp, i32, vp, vp, vp]
    L.hu_align_batch.argtypes = [vp, C.c_int]
    L.hu_batch_set_seeds.argtypes = [vp, vp, vp]
    L.hu_batch_set_estimates.argtypes = [vp, vp, vp]
    L.hu_batch_set_places.argtypes = [vp, vp, vp]
    for f in ("hu_seed_batch", "hu_estimate_batch", "hu_filter_batch", "hu_place_batch", "hu_finalize_batch"):
        getattr(L, f).argtypes = [vp, P(SepParams)]
    L.hu_assign_batch.argtypes = [vp, C.c_int, P(SepParams)]
    L.hu_batch_sync.argtypes = [vp]
    L.hu_batch_get_align.argtypes = [vp, vp, vp, i64, vp]
    L.hu_batch_get_merged.argtypes = [vp, vp]
    L.hu_batch_get_aligned_codes.argtypes = [vp, vp]
    L.hu_batch_get_seeds.argtypes = [vp, vp, vp, vp]
    L.hu_batch_get_estimates.argtypes = [vp, vp, vp]
    L.hu_batch_get_places.argtypes = [vp, vp, vp]
    L.hu_batch_get_best.argtypes = [vp, vp]
    L.hu_batch_stage_ms.argtypes = [vp, vp]
    L.hu_batch_get_stats.argtypes = [vp, P(BatchStats)]
    L.hu_trace_to_align.argtypes = [P(HmmParams), C.c_char_p, i32, P(AlignRow), C.c_char_p, vp]
    L.hu_host_alloc.restype = vp
    L.hu_host_alloc.argtypes = [C.c_size_t]
    L.hu_host_free.argtypes = [vp]
    L.hu_batch_mark.argtypes = [vp, C.c_int]
    L.hu_batch_elapsed.argtypes = [vp, C.c_int, vp, C.c_int, P(C.c_float)]
    L.hu_batch_get_seed_keys.argtypes = [vp, i32, vp, P(i32)]
    L.hu_debug_literal_order.argtypes = [C.c_int, vp, i32, i32, i32, vp, vp, vp]
    L.hu_debug_stdsort.argtypes = [vp, i32, vp]
    L.hu_batch_pack_rows.argtypes = [vp, i64, vp]
    L.hu_batch_stream.restype = vp
    L.hu_batch_stream.argtypes = [vp]
    _LIB = L
    return L


class HuError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        raise HuError("hmmufotu_b200 error %d: %s" % (rc, load().hu_last_error().decode()))
