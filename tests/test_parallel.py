"""World-size-2 gloo tests (CPU) of the multi-GPU plumbing: sample sharding, complex broadcast, design gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from abx_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, num_samples, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # rank 0 owns the complex, the other ranks start from zeros of the same shape
        g = torch.Generator().manual_seed(0)
        ref = {'seq': torch.randint(0, 20, (1, 17), generator=g), 'pos': torch.randn(1, 17, 14, 3, generator=g)}
        batch = {k: (v.clone() if rank == 0 else torch.zeros_like(v)) for k, v in ref.items()}
        batch['name'] = ['6ct7_H_L_S'] if rank == 0 else None
        batch = parallel.broadcast_complex(batch, ['seq', 'pos'])
        ok = all(torch.equal(batch[k], ref[k]) for k in ref) and batch['name'] == ['6ct7_H_L_S']

        mine = parallel.shard_samples(num_samples, rank, world)
        # "design" of sample k = a tensor that encodes k, so the gathered order can be checked
        local = {'atom14': torch.stack([torch.full((5, 3), float(k)) for k in mine]) if mine else torch.zeros(0, 5, 3),
                 'seq': torch.tensor([[k] * 4 for k in mine], dtype=torch.int64).reshape(len(mine), 4)}
        full = parallel.gather_designs(local, num_samples)
        if rank == 0:
            ok = ok and full['atom14'].shape == (num_samples, 5, 3) and \
                torch.equal(full['atom14'][:, 0, 0], torch.arange(num_samples, dtype=torch.float32)) and \
                torch.equal(full['seq'][:, 0], torch.arange(num_samples))
        else:
            ok = ok and full is None
        slow = parallel.max_over_ranks(10.0 + rank, torch.device('cpu'))
        ok = ok and slow == 10.0 + world - 1
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _run(world, num_samples):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_samples, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == {r: True for r in range(world)}


def test_shard_samples_partitions_every_index_once():
    for world in (1, 2, 4, 8):
        for n in (0, 1, 7, 16, 100, 1024):
            got = sorted(k for r in range(world) for k in parallel.shard_samples(n, r, world))
            assert got == list(range(n))
            sizes = [len(parallel.shard_samples(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_broadcast_and_gather_world2_even():
    _run(2, 8)


def test_broadcast_and_gather_world2_ragged():
    _run(2, 5)


def test_single_process_is_identity():
    local = {'x': torch.arange(6.).reshape(3, 2)}
    assert parallel.gather_designs(local, 3) is local
    b = {'seq': torch.zeros(1, 3)}
    assert parallel.broadcast_complex(b, ['seq']) is b
