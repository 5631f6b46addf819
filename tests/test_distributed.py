"""CPU, world_size 2 over gloo: row/star sharding and the result gather used by the
multi-GPU path (no GPU needed: the evaluator is a stand-in for lnlike_batch)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from minesweeper_b200 import distributed as D


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 9, 1000, 1 << 20):
        for size in (1, 2, 3, 4, 8):
            blocks = [D.shard_bounds(n, r, size) for r in range(size)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for (a, b), (c, d) in zip(blocks, blocks[1:]):
                assert b == c and a <= b and c <= d
            assert max(b - a for a, b in blocks) == (-(-n // size) if n else 0)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, size, port, n, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=size)
    try:
        rng = np.random.default_rng(0)
        theta = torch.from_numpy(rng.standard_normal((n, 6)))
        fn = lambda th: (th ** 2).sum(dim=1) * -0.5          # stand-in for the batched lnL
        full = D.sharded_lnlike(fn, theta)
        want = fn(theta)
        ok = bool(torch.equal(full, want))
        stars = D.star_blocks(11)
        der = torch.arange(n * 3, dtype=torch.float64).reshape(n, 3)
        lo, hi = D.shard_bounds(n, rank, size)
        g = D.gather_rows(der[lo:hi], n)
        ok = ok and bool(torch.equal(g, der))
        q.put((rank, ok, stars.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n', [10, 7, 1])
def test_sharded_lnlike_gloo_world2(n):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert res[0][2] + res[1][2] == list(range(11))      # stars partitioned, bit-identical union


def test_bind_to_gpu_numa_degrades_gracefully():
    """Without NVML (this container) the affinity helper reports None and leaves the process alone."""
    import os
    from minesweeper_b200.distributed import bind_to_gpu_numa
    before = os.sched_getaffinity(0)
    cores = bind_to_gpu_numa(0)
    assert cores is None or set(cores) <= before
    if cores is None:
        assert os.sched_getaffinity(0) == before
