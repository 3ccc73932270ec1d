"""CPU, world_size 2 over gloo: contiguous read sharding and the row gather keep the input order."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from hmmufotu_b200 import driver, synth


def test_shard_bounds_cover_and_order():
    for n in (0, 1, 7, 8, 1000):
        for w in (1, 2, 3, 8):
            cuts = [driver.shard_bounds(n, w, r) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def _fake_process(sub):
    """Stands in for the GPU: rows that depend only on the read itself."""
    from hmmufotu_b200 import engine
    a = np.zeros(sub.n, dtype=engine.ALIGN_DT)
    b = np.zeros(sub.n, dtype=engine.PLACE_DT)
    lens = np.diff(sub.offsets)
    a["seqEnd"] = lens
    a["cost"] = [float(sub.codes[sub.offsets[i]:sub.offsets[i + 1]].sum()) for i in range(sub.n)]
    b["cNode"] = sub.src_leaf
    b["nPlaces"] = 1
    return a, b


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    db = synth.make_db(n_leaves=10, cs_len=200, occupancy=0.9, seed=5)
    reads = synth.make_reads(db, 37, read_len=80, seed=2)
    rows = driver.assign_reads(reads, _fake_process, world, rank)
    if rank == 0:
        q.put(rows)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_process():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    rows2 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    db = synth.make_db(n_leaves=10, cs_len=200, occupancy=0.9, seed=5)
    reads = synth.make_reads(db, 37, read_len=80, seed=2)
    rows1 = driver.assign_reads(reads, _fake_process, 1, 0)
    assert rows2.shape == rows1.shape == (37,)
    assert np.array_equal(rows2["read"], np.arange(37))
    assert rows2.tobytes() == rows1.tobytes()
