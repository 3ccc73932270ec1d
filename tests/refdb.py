"""Readers for the reference's database files and assignment TSV (test helpers).

The formats are the ones written by the reference's hmmufotu-build / hmmufotu (SURVEY.md App. C;
/root/reference/src/PhyloTreeUnrooted.cpp:537-567, src/hmmufotu.cpp:586-600,736-746)."""
from __future__ import annotations

import gzip
import struct
from types import SimpleNamespace

import numpy as np

MODEL_TYPES = {"GTR": 0, "TN93": 1, "HKY85": 2, "F81": 3, "K80": 4, "JC69": 5}


def _open(path):
    return gzip.open(path, "rb") if path.endswith(".gz") else open(path, "rb")


def load_ptu(path: str) -> SimpleNamespace:
    """<db>.ptu -> namespace with the fields oracle_py.OracleTree / engine.Database take (+ up/dn messages, names)."""
    b = _open(path).read()
    assert b[:8] == b"HmmUFOtu", "not a HmmUFOtu binary file"
    o = 20
    n_nodes, = struct.unpack_from("<Q", b, o); o += 8
    cs_len, = struct.unpack_from("<i", b, o); o += 4

    def rstr():
        nonlocal o
        n, = struct.unpack_from("<Q", b, o); o += 8
        s = b[o:o + n]; o += n
        return s

    name, seq_name, anno = [None] * n_nodes, [None] * n_nodes, [None] * n_nodes
    anno_dist = np.zeros(n_nodes)
    seq = np.full((n_nodes, cs_len), -2, dtype=np.int8)
    for _ in range(n_nodes):
        i, = struct.unpack_from("<q", b, o); o += 8
        name[i] = rstr().decode()
        if b[o]:
            o += 1; rstr()
        else:
            o += 1
        seq_name[i] = rstr().decode()
        codes = rstr()
        if codes:
            seq[i] = np.frombuffer(codes, dtype=np.int8)
        anno[i] = rstr().decode()
        anno_dist[i], = struct.unpack_from("<d", b, o); o += 8
    n_edges, = struct.unpack_from("<Q", b, o); o += 8
    parent = np.full(n_nodes, -1, dtype=np.int32)
    blen = np.zeros(n_nodes)
    per = 4 * cs_len
    up = np.zeros((n_nodes, cs_len, 4)); dn = np.zeros((n_nodes, cs_len, 4))
    for _ in range(n_edges):
        id1, id2, is_par, length, N = struct.unpack_from("<qq?dQ", b, o); o += 8 + 8 + 1 + 8 + 8
        assert N == per
        msg = np.frombuffer(b, dtype="<f8", count=per, offset=o).reshape(cs_len, 4); o += per * 8
        if is_par:      # id1 is the parent of id2: message parent -> child
            parent[id2] = id1; blen[id2] = length; dn[id2] = msg
        else:           # message child -> parent
            up[id1] = msg
    root, = struct.unpack_from("<q", b, o); o += 8
    root_msg = np.frombuffer(b, dtype="<f8", count=per, offset=o).reshape(cs_len, 4).copy(); o += per * 8
    height = np.zeros(n_nodes)
    for _ in range(n_nodes):
        i, h = struct.unpack_from("<qd", b, o); o += 16
        height[i] = h
    n_idx, = struct.unpack_from("<I", b, o); o += 4
    o += n_idx * 12
    # text model, then the discrete-Gamma block
    end = b.index(b"\n", o)
    mtype = b[o:end].decode().strip(); o = end + 1
    lines = []
    n_model_lines = {"GTR": 13, "TN93": 4, "HKY85": 4, "F81": 4, "K80": 3, "JC69": 2}[mtype]
    for _ in range(n_model_lines):
        end = b.index(b"\n", o)
        lines.append(b[o:end].decode()); o = end + 1
    model_text = "\n".join(lines) + "\n"
    has_dg = b[o]; o += 1
    dg_rates = np.zeros(0)
    dg = None
    if has_dg:
        K, alpha = struct.unpack_from("<id", b, o); o += 12
        bnd = np.frombuffer(b, dtype="<f8", count=K + 1, offset=o); o += (K + 1) * 8
        dg_rates = np.frombuffer(b, dtype="<f8", count=K, offset=o).copy(); o += K * 8
        dg = dict(K=K, alpha=alpha, b=bnd.copy(), r=dg_rates)
    assert o == len(b), "trailing bytes in .ptu (%d of %d parsed)" % (o, len(b))
    # equal annotation strings share a taxon id (what the C ABI's hu_tree_desc.taxon means)
    ids = {}
    taxon = np.array([ids.setdefault(a, len(ids)) for a in anno], dtype=np.int32)
    model = parse_model(model_text)
    return SimpleNamespace(n_nodes=n_nodes, cs_len=cs_len, root=int(root), parent=parent, blen=blen, height=height, seq=seq,
                           taxon=taxon, anno=anno, anno_dist=anno_dist, name=name, seq_name=seq_name, up=up, dn=dn,
                           root_msg=root_msg, model=model, model_text=model_text, dg_rates=dg_rates, dg=dg)


def parse_model(text: str) -> dict:
    m = dict(type=0, pi=np.full(4, 0.25), R=np.zeros((4, 4)), kr=0.0, ky=0.0, kappa=0.0, beta=0.0)
    toks = [ln.split() for ln in text.splitlines() if ln and not ln.startswith("#")]
    i = 0
    while i < len(toks):
        t = toks[i]
        if t[0] == "Type:":
            m["type"] = MODEL_TYPES[t[1]]
        elif t[0] == "pi:":
            m["pi"] = np.array([float(x) for x in t[1:5]])
        elif t[0] == "R:":
            m["R"] = np.array([[float(x) for x in toks[i + 1 + r]] for r in range(4)])
            i += 4
        elif t[0] == "Q:":
            i += 4
        else:
            for k in range(0, len(t) - 1, 2):
                m[t[k].rstrip(":")] = float(t[k + 1])
        i += 1
    if m["type"] == MODEL_TYPES["K80"]:
        m["beta"] = 1.0 / (2.0 * m["kappa"])  # src/K80.h: setBeta()
    return m


def load_msa_cslen(path: str) -> int:
    b = _open(path).read(4096)
    o = 20
    for _ in range(2):
        n, = struct.unpack_from("<Q", b, o); o += 8 + n
    return struct.unpack_from("<II", b, o)[1]


def load_tsv(path: str):
    """assignment TSV -> (header list, {read id: row list})"""
    hdr, rows = None, {}
    for ln in _open(path).read().decode().splitlines():
        if ln.startswith("#") or not ln:
            continue
        f = ln.split("\t")
        if hdr is None:
            hdr = f
        else:
            rows[f[0]] = f
    return hdr, rows


def read_fasta(path: str):
    out, cur = [], None
    for ln in _open(path).read().decode().splitlines():
        if ln.startswith(">"):
            head = ln[1:].split(None, 1)
            cur = [head[0], head[1] if len(head) > 1 else "", []]
            out.append(cur)
        elif cur is not None:
            cur[2].append(ln.strip())
    return [(i, d, "".join(s)) for i, d, s in out]
