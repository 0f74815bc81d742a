"""Multi-GPU host logic on CPU: world_size-2 and -3 gloo runs of the galaxy sharding + ellipticity gather
(gdeconv/shard.py).  The per-stamp computation is replaced by a deterministic function of the galaxy index -- the
property under test is that every galaxy lands exactly once, in index order, for even, ragged and tiny totals."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _fake_e(lo, hi):
    i = torch.arange(lo, hi, dtype=torch.float32)
    return torch.stack([torch.sin(i), torch.cos(0.5 * i)], dim=1)


def _worker(rank, world, port, totals, q):
    import sys
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(here, 'galaxy-deconv_b200')]
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from gdeconv.shard import run_sharded, shard_range
    ok = True
    for n in totals:
        got = run_sharded(_fake_e, n)
        ok &= bool(torch.equal(got, _fake_e(0, n)))
        lo, hi = shard_range(n, rank, world)
        ok &= 0 <= lo <= hi <= n
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.parametrize('world', [2, 3])
def test_shard_and_gather(world):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, [10, 7, 1, 2, 1000], q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(30) for p in procs]
    assert sorted(r for r, _ in res) == list(range(world)) and all(ok for _, ok in res)


def test_shard_range_covers_everything():
    import sys
    from conftest import PKG
    from gdeconv.shard import shard_range
    for n in (0, 1, 5, 8, 1000003):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
