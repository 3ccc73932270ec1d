"""K2 on the tensor cores (hu_mm.cuh / hu_seedmm.cu) against the bit-plane scan it replaces: the seed keys of every (read,
node) must be bit-identical, and so must the seeds chosen from them; the standalone tile-kernel check of tests/cuda is
run too when nvcc is present."""
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

pytestmark = pytest.mark.gpu


def _keys_and_seeds(gdb, reads, max_len, paired=False, **sep):
    from hmmufotu_b200 import engine
    b = engine.Batch(gdb, reads.n, max_len, paired=paired)
    b.upload(reads)
    b.alignSeq()
    p = engine.SepParams(**sep)
    b.getSeed(p)
    keys = np.stack([b.seed_keys(i) for i in range(reads.n)])
    n, ids, _ = b.seeds()
    b.close()
    return keys, [ids[i, :n[i]] for i in range(reads.n)]


@pytest.mark.parametrize("shape", ["amplicon", "gapped_inserts", "long"])
def test_tensor_core_scan_equals_the_bit_plane_scan(shape, monkeypatch):
    from hmmufotu_b200 import _lib, engine, synth
    if shape == "amplicon":
        db = synth.make_db(n_leaves=700, cs_len=1500, occupancy=0.6, seed=11)
        reads = synth.make_reads(db, 300, read_len=250, seed=5)
        max_len = 280
    elif shape == "gapped_inserts":       # sparse alignment: many read bases land in non-match columns (the fix-up pass)
        db = synth.make_db(n_leaves=300, cs_len=2400, occupancy=0.35, seed=12)
        reads = synth.make_reads(db, 520, read_len=200, seed=6)
        max_len = 240
    else:
        db = synth.make_db(n_leaves=200, cs_len=1800, occupancy=0.8, seed=13)
        reads = synth.make_reads(db, 70, read_len=1100, seed=7)
        max_len = 1300
    gdb = engine.Database(db, device=0)
    monkeypatch.setenv("HU_SEED_SCAN", "popc")
    k0, s0 = _keys_and_seeds(gdb, reads, max_len)
    monkeypatch.setenv("HU_SEED_SCAN", "mma")
    k1, s1 = _keys_and_seeds(gdb, reads, max_len)
    # option filter in the (dist, id) order: excluded positions
    k2, s2 = _keys_and_seeds(gdb, reads, max_len, seed_order=_lib.SEED_ORDER_ID, max_height=float(np.median(db.height)))
    monkeypatch.setenv("HU_SEED_SCAN", "popc")
    k3, s3 = _keys_and_seeds(gdb, reads, max_len, seed_order=_lib.SEED_ORDER_ID, max_height=float(np.median(db.height)))
    gdb.close()
    assert k0.shape == k1.shape and (k0 != 0xFFFFFFFF).any()
    bad = np.argwhere(k0 != k1)
    assert bad.size == 0, "keys differ at (read, position) %s: popc %s mma %s" % (bad[:5].tolist(), k0[tuple(bad[:5].T)], k1[tuple(bad[:5].T)])
    assert all(np.array_equal(a, b) for a, b in zip(s0, s1))
    assert np.array_equal(k2, k3) and (k2 == 0xFFFFFFFF).any()
    assert all(np.array_equal(a, b) for a, b in zip(s2, s3))


def test_tile_kernel_against_a_host_gemm(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc on this box")
    exe = str(tmp_path / "mm_test")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "hmmufotu_b200", "csrc"),
                           os.path.join(HERE, "cuda", "mm_test.cu"), "-o", exe])
    for shape in (("3", "2", "7"), ("5", "3", "4"), ("2", "1", "1")):
        p = subprocess.run([exe, *shape], capture_output=True, text=True, timeout=120)
        assert p.returncode == 0 and "mismatched 0" in p.stdout, p.stdout + p.stderr
