"""Read-sharding driver: the database is replicated on every GPU, reads are split into contiguous blocks in
file order (so concatenating the ranks' rows reproduces the order the single-threaded reference writes,
/root/reference/src/hmmufotu.cpp:603-751), and the only collective is the gather of fixed-size assignment
rows to rank 0.  One process per GPU; `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is plumbing.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np

ROW_DT = np.dtype([("read", "<i8"), ("seqStart", "<i4"), ("seqEnd", "<i4"), ("hmmStart", "<i4"), ("hmmEnd", "<i4"),
                   ("csStart", "<i4"), ("csEnd", "<i4"), ("flags", "<i4"), ("nPlaces", "<i4"), ("cost", "<f8"),
                   ("cNode", "<i4"), ("pNode", "<i4"), ("aNode", "<i4"), ("taxon", "<i4"), ("ratio", "<f8"),
                   ("annoDist", "<f8"), ("loglik", "<f8"), ("qPlace", "<f8"), ("qTaxon", "<f8")])


def shard_bounds(n_reads: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; block sizes differ by at most one, earlier ranks get the extra."""
    base, extra = divmod(n_reads, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_rows(read0: int, align_rows: np.ndarray, best_rows: np.ndarray) -> np.ndarray:
    """Fixed-size POD rows (one per read) from the device results of one shard."""
    out = np.zeros(align_rows.shape[0], dtype=ROW_DT)
    out["read"] = read0 + np.arange(align_rows.shape[0])
    for f in ("seqStart", "seqEnd", "hmmStart", "hmmEnd", "csStart", "csEnd", "flags", "cost"):
        out[f] = align_rows[f]
    for f in ("nPlaces", "cNode", "pNode", "aNode", "taxon", "ratio", "annoDist", "loglik", "qPlace", "qTaxon"):
        out[f] = best_rows[f]
    return out


def gather_rows(rows: np.ndarray, n_total: int, world: int, rank: int, device: Optional[str] = None) -> Optional[np.ndarray]:
    """Gather every rank's rows on rank 0 in read order.  Equal-size padded blocks -> one all_gather."""
    if world == 1:
        return rows
    import torch
    import torch.distributed as dist
    per = -(-n_total // world)
    buf = np.zeros(per, dtype=ROW_DT)
    buf[:rows.shape[0]] = rows
    t = torch.from_numpy(buf.view(np.uint8).copy())
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    if rank != 0:
        return None
    out = []
    for r, p in enumerate(parts):
        lo, hi = shard_bounds(n_total, world, r)
        out.append(p.cpu().numpy().view(ROW_DT)[:hi - lo])
    return np.concatenate(out)


def assign_reads(reads, process_shard: Callable, world: int = 1, rank: int = 0, device: Optional[str] = None):
    """Split `reads` over ranks, run `process_shard(sub_reads) -> (align_rows, best_rows)` on this rank's block
    and gather the rows on rank 0 (None elsewhere)."""
    lo, hi = shard_bounds(reads.n, world, rank)
    sub = reads.slice(lo, hi)
    a, b = process_shard(sub)
    return gather_rows(pack_rows(lo, a, b), reads.n, world, rank, device)


class DeviceGather:
    """The gather of the path on device buffers: every rank's batch writes its fixed-size rows into a device tensor with a
    kernel on the batch stream (hu_batch_pack_rows); one NCCL all_gather, ordered after that kernel and asynchronous to the
    host, collects the ranks' blocks in rank (= file) order, so it overlaps the next batch's kernels on the other workspace.
    torch.distributed is the plumbing: the tensors are raw byte buffers and the batch's cudaStream_t is wrapped, not replaced."""

    def __init__(self, batch_rows: int, world: int, device: int):
        import torch
        self.torch, self.world, self.n = torch, world, batch_rows
        dev = torch.device("cuda", device)
        self.mine = torch.empty(batch_rows * ROW_DT.itemsize, dtype=torch.uint8, device=dev)
        self.all = torch.empty(world * batch_rows * ROW_DT.itemsize, dtype=torch.uint8, device=dev) if world > 1 else self.mine
        self.work = None

    def launch(self, batch, read0: int):
        """Enqueue: pack kernel on the batch stream, then (world > 1) the all_gather after it.  Returns at once."""
        import torch.distributed as dist
        torch = self.torch
        if self.world == 1:
            batch.pack_rows(read0, self.mine.data_ptr())
            return
        self.ext = torch.cuda.ExternalStream(batch.stream)
        with torch.cuda.stream(self.ext):
            if self.work is not None:
                self.work.wait()        # stream order: the previous gather has read `mine` before the pack kernel rewrites it
            batch.pack_rows(read0, self.mine.data_ptr())
            self.work = dist.all_gather_into_tensor(self.all, self.mine, async_op=True)

    def finish_on_stream(self):
        """Make the batch stream wait for the pending gather (so that an event recorded on it afterwards covers the gather)."""
        if self.work is not None:
            with self.torch.cuda.stream(self.ext):
                self.work.wait()
            self.work = None

    def wait(self):
        if self.work is not None:
            self.work.wait()
            self.work = None

    def rows(self) -> np.ndarray:
        """Host copy of the gathered rows (rank order = read order); synchronises."""
        self.wait()
        return self.all.cpu().numpy().view(ROW_DT)
