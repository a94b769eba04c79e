"""N > 1 host-side logic on CPU: world_size-2 gloo (the GPU collectives themselves are exercised by
tests/test_gpu_sharded.py on >= 2 GPUs)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from lasso_b200 import parallel

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ident = bytes(range(128)) if rank == 0 else b""
    got = parallel.broadcast_bytes(ident, src=0)  # how the job id of a sharded proof travels
    # the real thing: rank 0 draws the id through the C-ABI (no GPU needed for that call), every rank must end up with it
    import ctypes

    import lasso_b200 as lb

    real = (ctypes.c_uint8 * 128)()
    if rank == 0:
        assert lb.lib().lasso_comm_unique_id(real) == 0
    real_got = parallel.broadcast_bytes(bytes(real) if rank == 0 else b"", src=0)
    allids = [None] * world
    dist.all_gather_object(allids, real_got)
    assert len(real_got) == 128 and all(x == allids[0] for x in allids) and any(b != 0 for b in real_got)
    mx = parallel.max_over_ranks([1.0 + rank, 5.0 - rank])  # timings: max over ranks
    # the partition rule: rank g holds X[i*G + g]; bound_poly_var_top pairs (i, i + n/2) stay on one rank
    n = 64
    X = np.arange(n, dtype=np.int64)
    mine = parallel.shard_low_bits(X, rank, world)
    half = n // 2
    local_pairs_ok = all((int(v) + half) % world == rank for v in mine[: len(mine) // 2])
    # emulate one sharded bind + the gather-then-add of per-rank partial sums with plain integers
    r = 7
    bound_local = mine[: len(mine) // 2] + r * (mine[len(mine) // 2:] - mine[: len(mine) // 2])
    parts = [None] * world
    dist.all_gather_object(parts, int(bound_local.sum()))
    full_bound = X[:half] + r * (X[half:] - X[:half])
    q.put((rank, got == bytes(range(128)), mx, local_pairs_ok, sum(parts) == int(full_bound.sum()),
           mine.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_helpers():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, bc_ok, mx, pairs_ok, sum_ok, mine in res:
        assert bc_ok and pairs_ok and sum_ok
        assert mx == [2.0, 5.0]
    from lasso_b200 import parallel

    back = parallel.unshard_low_bits([np.array(res[0][5]), np.array(res[1][5])])
    assert back.tolist() == list(range(64))
