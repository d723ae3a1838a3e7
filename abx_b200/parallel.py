"""Multi-GPU plumbing of the sampler: one process per GPU, independent (complex x sample) instances per rank.

The reference only has a non-functional `mp.spawn` skeleton (inference.py:59-76,389-392: process group
init/destroy, no collective, every rank would redo all samples).  Here sample k of a complex goes to rank
`k mod world`; the static features of the complex are broadcast from rank 0 once and the designed
coordinates / sequences / pLDDT are gathered back.  Nothing is exchanged inside the sampling loop, so the
backend only matters for these two calls (NCCL over NVLink on the B200 box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_samples(num_samples, rank, world):
    """Indices of the samples rank `rank` designs (round-robin)."""
    return list(range(rank, num_samples, world))


def chunk_seed(base_seed, sample_indices):
    """Seed of the generator one batch of samples draws from: a stable hash of (base seed, the batch's sample
    indices).  Different base seeds or different batches never share a seed; the draws of sample k depend on
    which samples it is batched with (i.e. on --samples_per_batch and on the number of ranks)."""
    import hashlib
    h = hashlib.sha256(repr((int(base_seed), tuple(int(k) for k in sample_indices))).encode()).digest()
    return int.from_bytes(h[:8], 'little') & 0x7FFFFFFFFFFFFFFF


def _active(group=None):
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def broadcast_complex(batch, fields, src=0, device=None, group=None):
    """Broadcast the tensor `fields` of a collated complex from rank `src` in place; other entries of the
    dict (strings, lists used by the PDB writer) go through `broadcast_object_list`.  Shapes must already
    agree on every rank (each rank collates the same synthetic or on-disk complex header)."""
    if not _active(group):
        return batch
    for k in fields:
        t = batch[k]
        buf = t.to(device) if device is not None else t
        buf = buf.contiguous()
        dist.broadcast(buf, src, group=group)
        batch[k] = buf if device is not None else t.copy_(buf)
    rest = [{k: v for k, v in batch.items() if not torch.is_tensor(v)}]
    dist.broadcast_object_list(rest, src, group=group)
    batch.update(rest[0])
    return batch


def gather_designs(local, num_samples, dst=0, group=None):
    """Gather per-sample result tensors to rank `dst`.

    `local`: dict name -> tensor [n_local, ...] holding this rank's samples in `shard_samples` order.
    Returns on `dst` a dict name -> tensor [num_samples, ...] in global sample order (None elsewhere).
    Ranks may hold different counts (num_samples not divisible by world): shards are padded to the longest."""
    if not _active(group):
        return local
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n_max = (num_samples + world - 1) // world
    out = {} if rank == dst else None
    for name, t in local.items():
        pad = torch.zeros((n_max,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, bufs, dst=dst, group=group)
        if rank == dst:
            full = torch.empty((num_samples,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            for r in range(world):
                idx = shard_samples(num_samples, r, world)
                full[idx] = bufs[r][:len(idx)]
            out[name] = full
    return out


def max_over_ranks(value, device, group=None):
    """Max of a python float over ranks (device-side timing is reported as the slowest rank's)."""
    t = torch.tensor([value], device=device, dtype=torch.float64)
    if _active(group):
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t)
