"""Data-parallel sharding of the likelihood path over the GPUs of one box.

The path has no exchange step: every likelihood evaluation and
every star is independent, so ranks take contiguous blocks of rows (or stars), keep
the model tables replicated, and the only collective is one gather of fixed-size
results per batch (``all_gather_into_tensor`` over NCCL/NVLink; gloo in the CPU
tests).  One process per GPU, launched with torchrun.
"""
import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n, rank, size):
    """Contiguous block [lo, hi) of ``n`` items for ``rank``: ceil(n / size) per rank,
    the tail ranks may be short or empty."""
    per = -(-n // size) if size > 0 else n
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def gather_rows(local, n_total, group=None):
    """Concatenate per-rank blocks (made with ``shard_bounds``) of a tensor whose first
    axis is the sharded one; every rank receives the full result.  Blocks are padded to
    the common block size so that one ``all_gather_into_tensor`` moves everything."""
    rank, size = world()
    if size == 1:
        return local
    per = -(-n_total // size)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((size * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    return out[:n_total]


def sharded_lnlike(fn, theta, group=None):
    """Evaluate ``fn(theta_block) -> lnL_block (tensor)`` on this rank's block of rows and
    gather the full lnL on every rank.  ``fn`` is e.g.
    ``lambda th: likeobj.lnlike_batch(th, want_derived=False)[0]``."""
    rank, size = world()
    n = len(theta)
    lo, hi = shard_bounds(n, rank, size)
    block = theta[lo:hi]
    if hi > lo:
        lnl = fn(block)
    else:
        lnl = None
    if size == 1:
        return lnl
    if lnl is None:
        dev = theta.device if torch.is_tensor(theta) else torch.device('cpu')
        lnl = torch.empty((0,), dtype=torch.float64, device=dev)
    return gather_rows(lnl, n, group=group)


class PeerGather(object):
    """The path's one collective -- the gather of lnL -- without a collective kernel.

    Every rank owns a gather buffer ``[world * rows_per_rank]`` f64 in symmetric memory (torch's
    ``_symmetric_memory``: the same allocation mapped into every process of the box over NVLink / NVSwitch).  The
    tensor-core photometry kernel is handed the peer-mapped addresses of all of them (``Engine.set_lnl_peers``) and its
    epilogue stores each lnL[p] at ``rank * rows_per_rank + p`` of EVERY rank's buffer: the gathered vector
    materialises on all GPUs while the tiles finish, no NCCL kernel competes with the persistent kernel for SMs and
    no copy runs afterwards.  ``complete(i)`` orders the ranks (symmetric-memory barrier on the current stream: one
    small CTA) so that buffer i is whole everywhere.  Two buffers by default so the barrier of step k can overlap
    the kernel of step k + 1.  Traffic: (world - 1) * rows_per_rank * 8 B out and in per rank and step."""

    def __init__(self, rows_per_rank, device, group=None, buffers=2):
        import torch.distributed._symmetric_memory as symm
        self.rank, self.size = world()
        if self.size > 8:
            raise ValueError('PeerGather supports up to 8 ranks (one box)')
        self.n = int(rows_per_rank)
        g = group if group is not None else dist.group.WORLD
        self.bufs, self.hdls = [], []
        for _ in range(buffers):
            t = symm.empty(self.size * self.n, dtype=torch.float64, device=device)
            self.hdls.append(symm.rendezvous(t, g))
            self.bufs.append(t)

    def arm(self, engine, i):
        """Point the engine's likelihood kernel at buffer i of every rank."""
        engine.set_lnl_peers([int(p) for p in self.hdls[i].buffer_ptrs], offset=self.rank * self.n)

    def complete(self, i):
        """All ranks have finished writing buffer i (enqueued on the current stream)."""
        self.hdls[i].barrier(channel=i)

    def gathered(self, i):
        return self.bufs[i]

    @property
    def bytes_per_rank_per_step(self):
        return (self.size - 1) * self.n * 8


def star_blocks(n_stars, rank=None, size=None):
    """Catalog mode: the block of star indices pinned to this rank's device."""
    r, s = world()
    rank = r if rank is None else rank
    size = s if size is None else size
    lo, hi = shard_bounds(n_stars, rank, size)
    return np.arange(lo, hi)


def bind_to_gpu_numa(device_index):
    """Pin this process (and the pinned host buffers it allocates afterwards) to the CPU cores NVML reports as
    local to the GPU.  With one process per GPU the host<->device copies of every rank then stay on their own
    socket.  Returns the core list, or None when NVML or the affinity call is unavailable."""
    import os
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(int(device_index))
        ncpu = os.cpu_count() or 1
        words = nv.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None
