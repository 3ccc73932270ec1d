"""Synthetic database and read generator (inputs only; no hot-path arithmetic lives here).

Produces the shapes SURVEY.md section 8(d) names: a random rooted binary tree, a gapped 16S-like MSA
evolved down it, a plan-7 profile over the occupied columns and Illumina-like reads with the
ViterbiAlignPath records the host-side seed stage would hand to the device
(BandedHMMP7::buildAlignPath, /root/reference/src/BandedHMMP7.cpp:894-941).

Node ids follow the reference's explicit-stack DFS (/root/reference/src/PhyloTreeUnrooted.cpp:131-155):
root = 0, ids assigned at pop time, so the last child is numbered first.
Heights follow /root/reference/src/PhyloTreeUnrooted.cpp:274-287 (minimum distance to a leaf below).
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Optional

import numpy as np

GTR, TN93, HKY85, F81, K80, JC69 = range(6)
MODEL_NAMES = ["GTR", "TN93", "HKY85", "F81", "K80", "JC69"]

# data/gg_97_otus_*.sm values
GG97_PI = np.array([0.25143203728312, 0.2339699010077697, 0.3192819122955631, 0.1953161494135471])
GG97_GTR_R = dict(AC=0.174135544187311, AG=0.4570007541933747, AT=0.2465029385001669,
                  CG=0.1531807413945161, CT=0.8350020396890323, GT=0.2002128896019161)


def gg97_model(mtype: int = GTR) -> Dict:
    """Substitution-model description with the shipped gg_97 parameters."""
    R = np.zeros((4, 4))
    r = GG97_GTR_R
    for (i, j, v) in ((0, 1, r["AC"]), (0, 2, r["AG"]), (0, 3, r["AT"]), (1, 2, r["CG"]), (1, 3, r["CT"]), (2, 3, r["GT"])):
        R[i, j] = R[j, i] = v
    m = dict(type=mtype, pi=GG97_PI.copy(), R=R, kr=1.0, ky=1.0, kappa=1.0, beta=1.0)
    if mtype == TN93:
        m.update(kr=0.79605931231778826, ky=0.81397612002227493, beta=1.4446586871838933)
    elif mtype == HKY85:
        m.update(kappa=1.6100354323400632, beta=1.1165074599193707)
    elif mtype == F81:
        m.update(beta=1.3477984847710103)
    elif mtype == K80:
        m.update(kappa=1.6100354323400632, pi=np.full(4, 0.25))
    elif mtype == JC69:
        m.update(pi=np.full(4, 0.25))
    return m


def model_text(m: Dict) -> str:
    """Text form as the reference writes it into .sm / .ptu (e.g. /root/reference/src/GTR.cpp:83-90)."""
    t = MODEL_NAMES[m["type"]]
    fmt = lambda v: repr(float(v))
    s = "# DNA Substitution Model\nType: %s\n" % t
    if m["type"] == JC69:
        return s
    if m["type"] != K80:
        s += "pi: " + " ".join(fmt(v) for v in m["pi"]) + "\n"
    if m["type"] == GTR:
        s += "R:\n" + "".join(" ".join(fmt(v) for v in row) + "\n" for row in m["R"])
        Q = _rate_matrix(m)
        s += "Q:\n" + "".join(" ".join(fmt(v) for v in row) + "\n" for row in Q)
    elif m["type"] == TN93:
        s += "kr: %s ky: %s beta: %s\n" % (fmt(m["kr"]), fmt(m["ky"]), fmt(m["beta"]))
    elif m["type"] == HKY85:
        s += "kappa: %s beta: %s\n" % (fmt(m["kappa"]), fmt(m["beta"]))
    elif m["type"] == F81:
        s += "beta: %s\n" % fmt(m["beta"])
    elif m["type"] == K80:
        s += "kappa: %s\n" % fmt(m["kappa"])
    return s


def _rate_matrix(m: Dict) -> np.ndarray:
    """A rate matrix with the model's stationary distribution; used only to SIMULATE sequences."""
    pi = np.asarray(m["pi"], dtype=np.float64)
    t = m["type"]
    if t == GTR:
        R = np.asarray(m["R"], dtype=np.float64)
    else:
        R = np.ones((4, 4))
        kr = m["kr"] if t == TN93 else (m["kappa"] if t in (HKY85, K80) else 1.0)
        ky = m["ky"] if t == TN93 else (m["kappa"] if t in (HKY85, K80) else 1.0)
        R[0, 2] = R[2, 0] = kr
        R[1, 3] = R[3, 1] = ky
    Q = R * pi[None, :]
    np.fill_diagonal(Q, 0.0)
    np.fill_diagonal(Q, -Q.sum(axis=1))
    return Q / -np.trace(Q)


def _transition_matrices(m: Dict, t: np.ndarray) -> np.ndarray:
    """P(t) for an array of branch lengths, shape (n, 4, 4); simulation only."""
    Q = _rate_matrix(m)
    lam, U = np.linalg.eig(Q)
    lam = lam.real
    U = U.real
    Ui = np.linalg.inv(U)
    e = np.exp(lam[None, :] * np.asarray(t)[:, None])
    P = np.einsum("ik,nk,kj->nij", U, e, Ui)
    P = np.clip(P, 0.0, None)
    return P / P.sum(axis=2, keepdims=True)


@dataclasses.dataclass
class SynthDB:
    n_nodes: int
    cs_len: int
    root: int
    parent: np.ndarray      # int32 [n_nodes], -1 for the root
    blen: np.ndarray        # f64   [n_nodes], length of edge (id, parent)
    height: np.ndarray      # f64   [n_nodes]
    seq: np.ndarray         # int8  [n_nodes, cs_len]; internal rows are placeholders until inferred
    is_leaf: np.ndarray     # bool  [n_nodes]
    taxon: np.ndarray       # int32 [n_nodes] annotation-group id
    model: Dict
    dg_rates: np.ndarray    # f64 [K] (empty: no discrete Gamma)
    # profile
    K: int
    eM: np.ndarray          # f64 [(K+1), 4]
    eI: np.ndarray          # f64 [(K+1), 4]
    T: np.ndarray           # f64 [(K+1), 3, 3]
    p2cs: np.ndarray        # int32 [K+1] (1-based consensus column of profile position k)
    cs2p: np.ndarray        # int32 [cs_len+1]
    children: Optional[List[List[int]]] = None

    @property
    def leaves(self) -> np.ndarray:
        return np.flatnonzero(self.is_leaf)


def discrete_gamma_rates(alpha: float, K: int) -> np.ndarray:
    """Mean rates of K equal-probability Gamma(alpha, 1/alpha) categories (Yang 1994).

    The reference stores its rates in the .ptu (DiscreteGammaModel::load,
    /root/reference/src/DiscreteGammaModel.cpp:57-72); the run path only reads them, so any positive
    rates are a valid database.  These are the textbook ones."""
    from scipy.stats import gamma as G
    from scipy.special import gammainc
    b = G.ppf(np.arange(K + 1) / K, alpha, scale=1.0 / alpha)
    b[-1] = np.inf
    cdf = lambda x: gammainc(alpha + 1.0, np.where(np.isinf(x), 1e300, x * alpha))
    return K * (cdf(b[1:]) - cdf(b[:-1]))


def _sig6(x: np.ndarray) -> np.ndarray:
    """Round to 6 significant digits, as costs read back from a .hmm text file would be."""
    x = np.asarray(x, dtype=np.float64)
    out = x.copy()
    fin = np.isfinite(x) & (x != 0)
    mag = np.floor(np.log10(np.abs(x[fin])))
    scale = 10.0 ** (5 - mag)
    out[fin] = np.round(x[fin] * scale) / scale
    # go through the decimal text to get the same doubles a parser produces
    out[fin] = np.array([float("%.6g" % v) for v in out[fin]]) if fin.sum() < 2_000_000 else out[fin]
    return out


def make_db(n_leaves: int = 125, cs_len: int = 1486, occupancy: float = 0.87, seed: int = 20260101,
            model: Optional[Dict] = None, dg_alpha: float = 0.0, dg_k: int = 4, mean_blen: float = 0.05,
            leaf_del: float = 0.02, leaf_ins: float = 0.001, n_taxa_depth: int = 6) -> SynthDB:
    rng = np.random.default_rng(seed)
    model = model or gg97_model(GTR)
    n_nodes = 2 * n_leaves - 1
    # ---- random topology by random joins (coalescent-like, depth O(log n)) ----
    left = np.full(n_nodes, -1, dtype=np.int64)
    right = np.full(n_nodes, -1, dtype=np.int64)
    active = list(range(n_leaves))
    nxt = n_leaves
    order = rng.random(n_nodes)
    for _ in range(n_leaves - 1):
        i = int(order[nxt - n_leaves] * len(active)) % len(active)
        a = active[i]
        active[i] = active[-1]
        active.pop()
        j = int(rng.integers(len(active)))
        b = active[j]
        left[nxt], right[nxt] = a, b
        active[j] = nxt
        nxt += 1
    troot = n_nodes - 1
    # ---- ids by explicit-stack DFS, id at pop time, children pushed in order ----
    new_id = np.full(n_nodes, -1, dtype=np.int64)
    stack = [troot]
    k = 0
    while stack:
        x = stack.pop()
        new_id[x] = k
        k += 1
        if left[x] >= 0:
            stack.append(left[x])
            stack.append(right[x])
    parent = np.full(n_nodes, -1, dtype=np.int32)
    children: List[List[int]] = [[] for _ in range(n_nodes)]
    for x in range(n_nodes):
        if left[x] >= 0:
            for c in (left[x], right[x]):
                parent[new_id[c]] = new_id[x]
                children[new_id[x]].append(int(new_id[c]))
    is_leaf = np.zeros(n_nodes, dtype=bool)
    is_leaf[new_id[:n_leaves]] = True
    blen = np.clip(rng.exponential(mean_blen, n_nodes), 1e-4, 0.6)
    blen[0] = 0.0
    # depth (ids are a pre-order: parent id < child id)
    depth = np.zeros(n_nodes, dtype=np.int32)
    for x in range(1, n_nodes):
        depth[x] = depth[parent[x]] + 1
    # heights: min distance to a leaf below
    height = np.where(is_leaf, 0.0, np.inf)
    for x in range(n_nodes - 1, 0, -1):
        p = parent[x]
        h = height[x] + blen[x]
        if h < height[p]:
            height[p] = h
    # taxonomy groups: the ancestor at depth n_taxa_depth (or the node itself above that)
    taxon = np.arange(n_nodes, dtype=np.int32)
    for x in range(1, n_nodes):
        if depth[x] > n_taxa_depth:
            taxon[x] = taxon[parent[x]]
    # ---- sequences on the occupied columns ----
    occ = rng.random(cs_len) < occupancy
    occ[0] = occ[-1] = True
    occ_cols = np.flatnonzero(occ)
    n_occ = occ_cols.size
    pi = np.asarray(model["pi"], dtype=np.float64)
    rates = discrete_gamma_rates(dg_alpha, dg_k) if dg_alpha > 0 else np.zeros(0)
    site_rate = rates[rng.integers(0, dg_k, n_occ)] if dg_alpha > 0 else np.ones(n_occ)
    # conserved / variable site classes make the column profile 16S-like
    site_rate = site_rate * np.where(rng.random(n_occ) < 0.6, 0.15, 1.6)
    states = np.zeros((n_nodes, n_occ), dtype=np.int8)
    states[0] = rng.choice(4, size=n_occ, p=pi)
    by_depth = np.argsort(depth, kind="stable")
    bounds = np.searchsorted(depth[by_depth], np.arange(depth.max() + 2))
    rate_classes = np.unique(site_rate)
    for d in range(1, depth.max() + 1):
        nodes = by_depth[bounds[d]:bounds[d + 1]]
        for c0 in range(0, nodes.size, 2048):
            nd = nodes[c0:c0 + 2048]
            pb = states[parent[nd]]
            u = rng.random(pb.shape, dtype=np.float32)
            out = np.zeros(pb.shape, dtype=np.int8)
            for rc in rate_classes:
                cols = np.flatnonzero(site_rate == rc)
                P = _transition_matrices(model, blen[nd] * rc)
                cum = np.cumsum(P, axis=2).astype(np.float32)
                th = cum[np.arange(nd.size)[:, None], pb[:, cols]]
                uu = u[:, cols]
                out[:, cols] = (uu > th[..., 0]).astype(np.int8) + (uu > th[..., 1]) + (uu > th[..., 2])
            states[nd] = out
    seq = np.full((n_nodes, cs_len), -2, dtype=np.int8)
    seq[:, occ_cols] = states
    del states
    leaf_ids = np.flatnonzero(is_leaf)
    # per-leaf deletions on occupied columns and sparse insertions elsewhere
    tmp = seq[leaf_ids]
    n_del = rng.binomial(leaf_ids.size * n_occ, leaf_del)
    tmp[rng.integers(0, leaf_ids.size, n_del), occ_cols[rng.integers(0, n_occ, n_del)]] = -2
    un_cols = np.flatnonzero(~occ)
    if un_cols.size and leaf_ins > 0:
        n_ins = rng.binomial(leaf_ids.size * un_cols.size, leaf_ins)
        tmp[rng.integers(0, leaf_ids.size, n_ins), un_cols[rng.integers(0, un_cols.size, n_ins)]] = \
            rng.integers(0, 4, n_ins).astype(np.int8)
    seq[leaf_ids] = tmp
    # internal rows: ungapped placeholder (to be replaced by the arg-max of the upward message)
    internal = np.flatnonzero(~is_leaf)
    seq[internal] = np.where(seq[internal] < 0, 2, seq[internal])
    # ---- profile over the match columns (occupancy >= 0.5 among leaves) ----
    leaf_seq = seq[leaf_ids]
    colocc = (leaf_seq >= 0).mean(axis=0)
    match_cols = np.flatnonzero(colocc >= 0.5)
    K = match_cols.size
    p2cs = np.zeros(K + 1, dtype=np.int32)
    p2cs[1:] = match_cols + 1
    cs2p = np.zeros(cs_len + 1, dtype=np.int32)
    cs2p[p2cs[1:]] = np.arange(1, K + 1)
    cs2p[p2cs[K] + 1:] = K
    counts = np.stack([(leaf_seq[:, match_cols] == b).sum(axis=0) for b in range(4)], axis=1).astype(np.float64)
    prob = (counts + 0.5) / (counts + 0.5).sum(axis=1, keepdims=True)
    eM = np.full((K + 1, 4), -math.log(0.25))
    eM[1:] = -np.log(prob)
    eI = np.full((K + 1, 4), -math.log(0.25))
    gapf = np.clip(1.0 - colocc[match_cols], 0.004, 0.3)          # deletion propensity at column k
    T = np.full((K + 1, 3, 3), np.inf)
    pMD = np.empty(K + 1)
    pMD[:K] = gapf                                                 # k -> D(k+1)
    pMD[K] = 0.0
    pMI = np.full(K + 1, 0.006)
    T[:, 0, 0] = -np.log(1.0 - pMI - pMD)
    T[:, 0, 1] = -np.log(pMI)
    with np.errstate(divide="ignore"):
        T[:, 0, 2] = -np.log(pMD)
    T[:, 1, 0] = -math.log(0.7)
    T[:, 1, 1] = -math.log(0.3)
    T[:, 2, 0] = -math.log(0.8)
    T[:, 2, 2] = -math.log(0.2)
    T[0, 2, 0], T[0, 2, 2] = 0.0, np.inf                           # BandedHMMP7.cpp:306-311 specials
    T[K, 2, 0], T[K, 2, 2] = 0.0, np.inf
    T[:, 1, 2] = np.inf
    T[:, 2, 1] = np.inf
    eM, eI, T = _sig6(eM), _sig6(eI), _sig6(T)
    return SynthDB(n_nodes=n_nodes, cs_len=cs_len, root=0, parent=parent, blen=blen, height=height,
                   seq=seq, is_leaf=is_leaf, taxon=taxon, model=model, dg_rates=rates, K=K, eM=eM, eI=eI, T=T,
                   p2cs=p2cs, cs2p=cs2p, children=children)


def write_hmm(db: SynthDB, path: str, name: str = "synth") -> None:
    """HMMER3/f-like text as BandedHMMP7 writes it (/root/reference/src/BandedHMMP7.cpp:324-378)."""
    def f(v):
        return "*" if not np.isfinite(v) else "%.6g" % v
    K = db.K
    with open(path, "w") as o:
        o.write("HMMER3/f\t1.5.1\nNAME\t%s\nLENG\t%d\nALPH\tDNA\nMAXL\t%d\nRF\tno\nMM\tno\nCONS\tyes\nCS\tno\nMAP\tyes\n" % (name, K, db.cs_len))
        o.write("HMM\tA\tC\tG\tT\n\tm->m\tm->i\tm->d\ti->m\ti->i\td->m\td->d\n")
        for k in range(K + 1):
            if k == 0:
                o.write("\tCOMPO\t" + "\t".join(f(v) for v in db.eM[0]) + "\n")
            else:
                o.write("\t%d\t" % k + "\t".join(f(v) for v in db.eM[k]) + "\t%d\t-\n" % db.p2cs[k])
            o.write("\t\t" + "\t".join(f(v) for v in db.eI[k]) + "\n")
            t = db.T[k]
            o.write("\t\t" + "\t".join(f(v) for v in (t[0, 0], t[0, 1], t[0, 2], t[1, 0], t[1, 1], t[2, 0], t[2, 2])) + "\n")
        o.write("//\n")


@dataclasses.dataclass
class ReadSet:
    codes: np.ndarray       # int8, concatenated base codes 0..3
    offsets: np.ndarray     # int64 [n+1]
    vpaths: np.ndarray      # int32 [n, 2, 6] {start,end,from,to,nIns,nDel}
    n_paths: np.ndarray     # int32 [n]
    src_leaf: np.ndarray    # int32 [n]

    @property
    def n(self) -> int:
        return int(self.n_paths.size)

    def read_str(self, i: int) -> str:
        return "".join("ACGT"[c] for c in self.codes[self.offsets[i]:self.offsets[i + 1]])

    def slice(self, lo: int, hi: int) -> "ReadSet":
        o = self.offsets[lo:hi + 1]
        return ReadSet(self.codes[o[0]:o[-1]].copy(), (o - o[0]).astype(np.int64), self.vpaths[lo:hi].copy(),
                       self.n_paths[lo:hi].copy(), self.src_leaf[lo:hi].copy())


def _build_vpath(db: SynthDB, cols: np.ndarray, frm: int) -> Optional[np.ndarray]:
    """buildAlignPath over the consensus columns of one exact seed hit (cols: 0-based column of each seed base)."""
    start = end = nins = ndel = 0
    seedcols = set(int(c) for c in cols)
    for j in range(int(cols[0]) + 1, int(cols[-1]) + 2):      # 1-based consensus columns csStart..csEnd
        k = int(db.cs2p[j])
        non_gap = (j - 1) in seedcols
        if k != 0:
            if start == 0:
                start = k
            end = k
            if not non_gap:
                ndel += 1
        elif non_gap:
            nins += 1
    to = frm + len(cols) - 1
    if not (start > 0 and start <= end and frm > 0 and frm <= to):
        return None
    return np.array([start, end, frm, to, nins, ndel], dtype=np.int32)


def _one_read(db: SynthDB, rng, row: np.ndarray, cols: np.ndarray, sub_rate: float, indel_prob: float, seed_len: int,
              seed_region: int, drop_seed: bool):
    """Copy the residues at `cols` of one leaf row, mutate lightly, and return (bases, vpaths) with the
    ViterbiAlignPaths of the first exact 5' and 3' seeds (HmmUFOtu_main.cpp:52-84)."""
    cols = cols.copy()
    L = cols.size
    bases = row[cols].copy()
    exact = np.ones(L, dtype=bool)                 # base i is an unmodified copy of column cols[i]
    mut = rng.random(L) < sub_rate
    bases[mut] = (bases[mut] + rng.integers(1, 4, int(mut.sum()))) % 4
    exact[mut] = False
    if rng.random() < indel_prob and L > 60:       # one deletion
        i = int(rng.integers(25, L - 25))
        bases = np.delete(bases, i); cols = np.delete(cols, i); exact = np.delete(exact, i)
        exact[i - 1:i + 1] = False
    if rng.random() < indel_prob and L > 60:       # one insertion
        i = int(rng.integers(25, bases.size - 25))
        bases = np.insert(bases, i, rng.integers(0, 4)); cols = np.insert(cols, i, -1); exact = np.insert(exact, i, False)
    N = bases.size
    paths = []
    region = min(seed_region, N)
    if not drop_seed:
        for sf in range(0, region - seed_len + 1):  # 5' seed
            if exact[sf:sf + seed_len].all():
                vp = _build_vpath(db, cols[sf:sf + seed_len], sf + 1)
                if vp is not None:
                    paths.append(vp)
                    break
        if not paths or N >= 2 * region:             # 3' seed
            for st in range(N - 1, N - region + seed_len - 2, -1):
                sf = st - seed_len + 1
                if sf < 0:
                    break
                if exact[sf:st + 1].all():
                    vp = _build_vpath(db, cols[sf:st + 1], sf + 1)
                    if vp is not None:
                        paths.append(vp)
                        break
    return bases.astype(np.int8), paths


def _pack(reads, src) -> ReadSet:
    n = len(reads)
    vps = np.zeros((n, 2, 6), dtype=np.int32)
    nps = np.zeros(n, dtype=np.int32)
    offsets = np.zeros(n + 1, dtype=np.int64)
    for r, (bases, paths) in enumerate(reads):
        nps[r] = len(paths)
        for i, vp in enumerate(paths):
            vps[r, i] = vp
        offsets[r + 1] = offsets[r] + bases.size
    return ReadSet(np.concatenate([b for b, _ in reads]), offsets, vps, nps, np.asarray(src, dtype=np.int32))


def make_reads(db: SynthDB, n: int, read_len: int = 250, seed: int = 1, window: Optional[tuple] = None,
               sub_rate: float = 0.01, indel_prob: float = 0.1, seed_len: int = 20, seed_region: int = 50,
               len_sd: float = 0.0, drop_seed_prob: float = 0.02) -> ReadSet:
    """Reads copied from random leaves, lightly mutated, with the VPaths an exact-match seed search finds.

    window=(c0, c1): every read starts at the first residue at or after a column drawn from [c0, c1)
    (the "V4-like" shape of SURVEY.md 8(d) config C3); None: anywhere the read fits."""
    rng = np.random.default_rng(seed)
    leaves = db.leaves
    reads, src = [], []
    for r in range(n):
        leaf = int(leaves[rng.integers(leaves.size)])
        row = db.seq[leaf]
        pos = np.flatnonzero(row >= 0)
        L = read_len if len_sd <= 0 else int(max(seed_len * 3, round(rng.normal(read_len, len_sd))))
        L = min(L, pos.size)
        if window is not None:
            c0 = int(rng.integers(window[0], window[1]))
            s = int(np.searchsorted(pos, c0))
            s = min(s, pos.size - L)
        else:
            s = int(rng.integers(0, pos.size - L + 1))
        drop = rng.random() < drop_seed_prob           # occasionally no seed at all -> full Viterbi
        reads.append(_one_read(db, rng, row, pos[s:s + L], sub_rate, indel_prob, seed_len, seed_region, drop))
        src.append(leaf)
    return _pack(reads, src)


def make_pairs(db: SynthDB, n: int, read_len: int = 150, insert_len: int = 260, insert_sd: float = 15.0, seed: int = 1,
               sub_rate: float = 0.01, indel_prob: float = 0.1, swap_prob: float = 0.0):
    """Paired-end reads: a fragment of ~insert_len residues of one leaf; the forward read is its first read_len
    bases, the mate its last read_len bases -- returned already in forward orientation, as the reference holds
    it after `revSeqI.nextSeq().revcom()` (src/hmmufotu.cpp:609).  swap_prob: fraction of pairs handed over in the
    wrong order (mate upstream), which the orientation test must flag as chimera.  -> (fwd ReadSet, mate ReadSet)"""
    rng = np.random.default_rng(seed)
    leaves = db.leaves
    fwd, rev, src = [], [], []
    for r in range(n):
        leaf = int(leaves[rng.integers(leaves.size)])
        row = db.seq[leaf]
        pos = np.flatnonzero(row >= 0)
        F = int(np.clip(round(rng.normal(insert_len, insert_sd)), read_len + 10, pos.size))
        s = int(rng.integers(0, pos.size - F + 1))
        a = _one_read(db, rng, row, pos[s:s + read_len], sub_rate, indel_prob, 20, 50, False)
        b = _one_read(db, rng, row, pos[s + F - read_len:s + F], sub_rate, indel_prob, 20, 50, False)
        if rng.random() < swap_prob:
            a, b = b, a
        fwd.append(a); rev.append(b); src.append(leaf)
    return _pack(fwd, src), _pack(rev, src)


# ---------------------------------------------------------------------------------------------------------
# Database files in the reference's formats: <db>.msa, <db>.hmm, <db>.ptu.
# (<db>.csfm holds libcds structures; this repo's host layer rebuilds the index from <db>.msa instead.)
# ---------------------------------------------------------------------------------------------------------
import struct

_MAGIC = b"HmmUFOtu" + struct.pack("<iii", 1, 5, 1)


def _wstr(f, s) -> None:
    b = s if isinstance(s, (bytes, bytearray)) else s.encode()
    f.write(struct.pack("<Q", len(b)))
    f.write(b)


def write_msa(db: SynthDB, path: str, name: str = "synth") -> None:
    """MSA::save layout (/root/reference/src/MSA.cpp:295-360); rows = leaves in id order."""
    leaves = db.leaves
    rows = db.seq[leaves]
    chars = np.where(rows >= 0, np.frombuffer(b"ACGT", dtype=np.uint8)[np.clip(rows, 0, 3)], ord("-")).astype(np.uint8)
    n, L = rows.shape
    with open(path, "wb") as f:
        f.write(_MAGIC)
        _wstr(f, "DNA")
        _wstr(f, name)
        f.write(struct.pack("<II", n, L))
        cs = np.full(L, ord("-"), dtype=np.uint8)
        for j in range(L):
            col = rows[:, j]
            col = col[col >= 0]
            if col.size * 2 >= n:
                cs[j] = b"ACGT"[int(np.bincount(col, minlength=4).argmax())]
        _wstr(f, cs.tobytes())
        f.write(struct.pack("<?", True))
        for x in leaves:
            _wstr(f, "leaf%d" % x)
        _wstr(f, chars.tobytes())
        valid = rows >= 0
        first = np.where(valid.any(axis=1), valid.argmax(axis=1), -1).astype(np.int32)
        last = np.where(valid.any(axis=1), L - 1 - valid[:, ::-1].argmax(axis=1), -1).astype(np.int32)
        f.write(first.tobytes()); f.write(last.tobytes()); f.write(valid.sum(axis=1).astype(np.int32).tobytes())
        res = np.stack([(rows == b).sum(axis=0) for b in range(4)], axis=1).astype(np.int32)      # column-major 4 x L
        gap = (~valid).sum(axis=0).astype(np.int32)
        f.write(res.tobytes()); f.write(gap.tobytes())
        f.write(np.ones(n).tobytes()); f.write(res.astype(np.float64).tobytes()); f.write(gap.astype(np.float64).tobytes())


def write_ptu(db: SynthDB, up: np.ndarray, dn: np.ndarray, path: str, anno=None) -> None:
    """PTUnrooted::save layout (/root/reference/src/PhyloTreeUnrooted.cpp:537-567).  up/dn: [node][col][4] log-domain
    messages; node sequences as currently held in db.seq (internal rows = inferred)."""
    n, L = db.n_nodes, db.cs_len
    anno = anno or ["k__Synth;p__Clade%d" % t for t in db.taxon]
    nbrs = [[] for _ in range(n)]
    for x in range(n):
        if db.parent[x] >= 0:
            nbrs[x].append(int(db.parent[x]))
    for x in range(n):                      # children after the parent, ascending id
        if db.parent[x] >= 0:
            nbrs[int(db.parent[x])].append(x)
    with open(path, "wb") as f:
        f.write(_MAGIC)
        f.write(struct.pack("<Qi", n, L))
        for x in range(n):
            f.write(struct.pack("<q", x))
            _wstr(f, "node%d" % x)
            f.write(struct.pack("<?", False))   # PTUNode::save writes the DigitalSeq without its alphabet
            _wstr(f, "node%d" % x)              # DigitalSeq name (the MSA row id in a real database)
            _wstr(f, db.seq[x].astype(np.int8).tobytes())
            _wstr(f, anno[x])
            f.write(struct.pack("<d", 0.0))
        f.write(struct.pack("<Q", 2 * (n - 1)))
        for u in range(n):
            for v in nbrs[u]:
                is_par = db.parent[v] == u
                f.write(struct.pack("<qq?d", u, v, bool(is_par), float(db.blen[v] if is_par else db.blen[u])))
                f.write(struct.pack("<Q", 4 * L))
                f.write(np.ascontiguousarray(dn[v] if is_par else up[u], dtype=np.float64).tobytes())
        f.write(struct.pack("<q", db.root))
        f.write(np.zeros(4 * L).tobytes())
        for x in range(n):
            f.write(struct.pack("<qd", x, float(db.height[x])))
        leaves = db.leaves
        f.write(struct.pack("<I", leaves.size))
        for i, x in enumerate(leaves):
            f.write(struct.pack("<Iq", i, int(x)))
        f.write((MODEL_NAMES[db.model["type"]] + "\n" + model_text(db.model)).encode())
        if db.dg_rates.size:
            K = db.dg_rates.size
            f.write(struct.pack("<?id", True, K, 0.5))
            f.write(np.zeros(K + 1).tobytes())
            f.write(np.asarray(db.dg_rates, dtype=np.float64).tobytes())
        else:
            f.write(struct.pack("<?", False))


def write_fasta(db_reads: "ReadSet", path: str, prefix: str = "read") -> None:
    with open(path, "w") as f:
        for i in range(db_reads.n):
            f.write(">%s%d src_leaf=%d\n%s\n" % (prefix, i, db_reads.src_leaf[i], db_reads.read_str(i)))


def write_db(db: SynthDB, up: np.ndarray, dn: np.ndarray, prefix: str) -> None:
    write_msa(db, prefix + ".msa")
    write_hmm(db, prefix + ".hmm")
    write_ptu(db, up, dn, prefix + ".ptu")
