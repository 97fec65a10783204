"""world_size-2 test (gloo, CPU) of the N > 1 host logic of bench.py: each rank owns a disjoint
slice of the point stream (no data-path collective), and the reported time is the MAX over
ranks, the throughput the SUM of the per-rank work over that time."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, npts, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "bsplinekit.jl_b200"))
    from bsplinekit_b200 import synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = synth.u_torch(5, rank * npts, npts, device="cpu")          # this rank's shard
    gathered = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(gathered, x)                                      # test-only: compare with the full stream
    ms = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)      # pretend per-rank device time
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    value = world * npts / (ms.item() * 1e-3) / 1e9
    if rank == 0:
        q.put((torch.cat(gathered).numpy(), ms.item(), value))
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_and_max_over_ranks():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "bsplinekit.jl_b200"))
    from bsplinekit_b200 import synth
    world, npts = 2, 1 << 12
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, npts, q)) for r in range(world)]
    for p in procs:
        p.start()
    allx, ms, value = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(allx, synth.u_numpy(5, 0, world * npts))    # disjoint, ordered shards
    assert ms == 15.0                                                  # max over ranks
    assert value == world * npts / 15e-3 / 1e9
