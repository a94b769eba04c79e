"""Host-side helpers for the multi-GPU modes (one process per GPU, torch.distributed for the plumbing).

Two modes:
  * independent proofs per GPU (bench.py default, weak scaling): no data-path collective at all;
  * ONE proof sharded over G GPUs (`Context.init_comm`): every polynomial is partitioned by the LOW log2(G) bits
    of its index — rank g holds X[i*G + g] — so bound_poly_var_top's pairs (i, i + n/2) are always local.
    The exchanges themselves (NCCL) live in csrc/comm.cu; the functions here state the partition rule and carry
    the few bytes that travel out of band.
"""
import numpy as np


def shard_low_bits(x, rank, world):
    """this rank's shard of a global array: elements rank, rank + world, rank + 2*world, ..."""
    x = np.asarray(x)
    assert x.shape[0] % world == 0
    return np.ascontiguousarray(x[rank::world])


def unshard_low_bits(shards):
    """inverse of shard_low_bits given the shards of all ranks in rank order"""
    world = len(shards)
    n = shards[0].shape[0]
    out = np.empty((n * world,) + shards[0].shape[1:], dtype=shards[0].dtype)
    for g, sh in enumerate(shards):
        out[g::world] = sh
    return out


def broadcast_bytes(data, src=0):
    """broadcast a bytes object (e.g. the 128-byte NCCL id) from `src` over torch.distributed (any backend)"""
    import torch
    import torch.distributed as dist

    n = torch.tensor([len(data) if dist.get_rank() == src else 0], dtype=torch.int64)
    cuda = dist.get_backend() == "nccl"
    if cuda:
        n = n.cuda()
    dist.broadcast(n, src=src)
    buf = torch.zeros(int(n.item()), dtype=torch.uint8)
    if dist.get_rank() == src:
        buf = torch.frombuffer(bytearray(data), dtype=torch.uint8).clone()
    if cuda:
        buf = buf.cuda()
    dist.broadcast(buf, src=src)
    return bytes(buf.cpu().numpy().tobytes())


def max_over_ranks(values):
    """element-wise max of a list of floats over all ranks (multi-GPU timings are the max over ranks)"""
    import torch
    import torch.distributed as dist

    t = torch.tensor(list(values), dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]
