"""Multi-GPU plumbing of the --zscore shuffle batch (one process per GPU).

The shuffles of the loop at reference src/ractip.cpp:1638-1657 are independent
(only four float accumulators cross iterations, :1626-1627,1655-1656), so rank r
of W takes shuffles r, r+W, r+2W, ... and runs the probability stage on them.
What the host-side ILP needs afterwards -- the thresholded variable lists x, y, z
and the accessible-region lists v, w of every shuffle (src/ractip.cpp:557-567,
578-588, 598-609, 619-628, 639-648; the lists carry every probability the integer
programme reads, so no dense table travels) -- is exchanged with ONE all-gather of
a fixed-capacity byte buffer per rank:

    [ rp_rec records | rp_sparse_counts ]      (each part 256-B aligned)

Capacities come from rp_sparse_plan (a bound, not a count), so every rank's
buffer has the same size and no size exchange is needed.  torch.distributed does
the transport (NCCL over NVLink on GPUs, gloo in the CPU tests); the kernels
write straight into the buffer through rp_batch_sparse_device.  Stream order:
rp_batch_sparse_device is asynchronous on the CONTEXT's stream.  gather() therefore
synchronises the batch that fill()ed the buffer before the collective, unless the
caller made the context's stream the one the collective runs on
(rp_set_stream(torch's current stream), as bench.py does) and says so with
same_stream=True.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import RpOpts, RpPair, RpSparseLayout
from .stage import REC_DTYPE, PairRecords

ALIGN = 256
CNT_BYTES = 24   # sizeof(rp_sparse_counts): n_x, n_y, n_z, overflow, n_v, n_w


def shard_indices(n: int, rank: int, world: int) -> List[int]:
    """Interleaved shard: balanced even if costs drift along the batch."""
    return list(range(rank, n, world))


def _plan(pairs: Sequence[Tuple[str, str]], opts: RpOpts):
    lib = _lib.load()
    n = len(pairs)
    arr = (RpPair * max(n, 1))()
    keep = []
    for k, (a, b) in enumerate(pairs):
        ba, bb = a.encode(), b.encode()
        keep.append((ba, bb))
        arr[k].s1, arr[k].n1, arr[k].s2, arr[k].n2 = ba, len(ba), bb, len(bb)
    lay = (RpSparseLayout * max(n, 1))()
    tr, tf = C.c_size_t(), C.c_size_t()
    rc = lib.rp_sparse_plan(arr, n, C.byref(opts), lay, C.byref(tr), C.byref(tf))
    if rc:
        raise RuntimeError("rp_sparse_plan failed: %d" % rc)
    return lay, tr.value, tf.value


def _up(x: int) -> int:
    return (x + ALIGN - 1) // ALIGN * ALIGN


class ShardPlan:
    """Who computes what, and the layout of the gathered buffer."""

    def __init__(self, pairs: Sequence[Tuple[str, str]], opts: RpOpts, rank: int, world: int):
        self.pairs, self.opts, self.rank, self.world = list(pairs), opts, rank, world
        self.n = len(self.pairs)
        self.shards = [shard_indices(self.n, r, world) for r in range(world)]
        self.layouts, recs, cnts = [], [], []
        for r in range(world):
            lay, tr, _tf = _plan([self.pairs[i] for i in self.shards[r]], opts)
            self.layouts.append(lay)
            recs.append(tr * REC_DTYPE.itemsize)
            cnts.append(len(self.shards[r]) * CNT_BYTES)
        # identical section sizes on every rank: the maximum over ranks
        self.rec_bytes, self.cnt_bytes = _up(max(recs)), _up(max(cnts))
        self.nbytes = self.rec_bytes + self.cnt_bytes

    @property
    def my_pairs(self) -> List[Tuple[str, str]]:
        return [self.pairs[i] for i in self.shards[self.rank]]

    def section_offsets(self) -> Tuple[int, int]:
        return 0, self.rec_bytes

    def capacity(self) -> int:
        """Records the local buffer can hold: what rp_batch_sparse_device is told."""
        return self.rec_bytes // REC_DTYPE.itemsize

    def fill(self, batch, local) -> None:
        """Have the batch's kernels write this rank's lists and counts into `local` (a CUDA uint8 tensor of
        nbytes): rp_batch_sparse_device without the dense window tables."""
        base = local.data_ptr()
        batch.stage._check(batch.lib.rp_batch_sparse_device(batch.handle, C.c_void_p(base), self.capacity(),
                                                            C.c_void_p(0), 0, C.c_void_p(base + self.rec_bytes)))
        self._filled_by = batch   # gather() waits for it unless told that the collective runs on the same stream

    def rec_view(self, buf: np.ndarray) -> np.ndarray:
        """The record section of one rank's buffer as a structured array."""
        o_rec = 0
        usable = self.rec_bytes // REC_DTYPE.itemsize * REC_DTYPE.itemsize
        return buf[o_rec:o_rec + usable].view(REC_DTYPE)

    def gather(self, local, group=None, same_stream: bool = False):
        """The single collective of the path.  `local` is a uint8 tensor of nbytes (CUDA or CPU).  The kernels that
        fill() launched run on the CONTEXT's stream: unless the caller has made that the stream the collective runs on
        (rp_set_stream with torch's current stream -- then pass same_stream=True), the batch is synchronised first."""
        import torch
        import torch.distributed as dist
        assert local.dtype == torch.uint8 and local.numel() == self.nbytes
        filled_by = getattr(self, "_filled_by", None)
        if filled_by is not None and not same_stream:
            filled_by.sync()
        self._filled_by = None
        out = torch.empty(self.world * self.nbytes, dtype=torch.uint8, device=local.device)
        if self.world == 1:
            out.copy_(local)
            return out
        dist.all_gather_into_tensor(out, local, group=group)
        return out

    def unpack(self, gathered: np.ndarray) -> List[PairRecords]:
        """Gathered bytes (host) -> per-shuffle records in the ORIGINAL batch order."""
        gathered = np.ascontiguousarray(gathered).view(np.uint8).reshape(self.world, self.nbytes)
        out: List[PairRecords] = [None] * self.n  # type: ignore
        o_rec, o_cnt = self.section_offsets()
        for r in range(self.world):
            buf = gathered[r]
            recs = self.rec_view(buf)
            cnts = buf[o_cnt:o_cnt + len(self.shards[r]) * CNT_BYTES].view(np.int32).reshape(-1, CNT_BYTES // 4)
            for k, gi in enumerate(self.shards[r]):
                S = self.layouts[r][k]
                nx, ny, nz, ov, nv, nw = (int(v) for v in cnts[k])
                if ov:
                    raise RuntimeError(f"record capacity exceeded for shuffle {gi}")
                out[gi] = PairRecords(
                    x=recs[S.x:S.x + nx], y=recs[S.y:S.y + ny], z=recs[S.z:S.z + nz], up1=None, up2=None,
                    v=recs[S.v:S.v + nv], w=recs[S.w:S.w + nw])
        return out
