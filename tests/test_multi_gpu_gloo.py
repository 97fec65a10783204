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


def _bench():
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_shard_points_partitions_the_stream():
    """bench.shard_points: the slices of all ranks are disjoint, contiguous, in rank order, of even length (the
    kernels work on 16-byte vectors); strong scaling splits ONE problem (BASELINE.json: 2^28 points "at 1/2/4/8
    B200"), weak scaling gives every rank the full count."""
    b = _bench()
    for points in (1 << 28, 1 << 20, 1000003, 10):
        for world in (1, 2, 3, 4, 8):
            for mode in ("strong", "weak"):
                slices = [b.shard_points(points, mode, world, r) for r in range(world)]
                npts = slices[0][0]
                assert all(s[0] == npts for s in slices) and npts % 2 == 0
                assert [s[1] for s in slices] == [r * npts for r in range(world)]
                assert all(s[2] == npts * world for s in slices)
                if mode == "strong":
                    assert points - 2 * world < npts * world <= points
                else:
                    assert npts == points // 2 * 2


def _strong_worker(rank, world, port, points, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "bsplinekit.jl_b200"))
    sys.path.insert(0, root)
    from bsplinekit_b200 import synth
    from oracle import oracle as O
    b = _bench()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    npts, first, total = b.shard_points(points, "strong", world, rank)
    # every rank evaluates ITS slice with a replicated basis (here: the CPU checker stands in for the device)
    k = 4
    t = O.augment_knots(synth.breakpoints(3, 50), k)
    c = 2 * synth.u_numpy(4, 0, t.size - k) - 1
    x = synth.u_numpy(5, first, npts)
    y = torch.from_numpy(O.spline_eval(0, k, t, c, x))
    gathered = [torch.empty_like(y) for _ in range(world)]
    dist.all_gather(gathered, y)                                     # test-only
    ms = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((torch.cat(gathered).numpy(), total, world * npts / (ms.item() * 1e-3) / 1e9))
    dist.barrier()
    dist.destroy_process_group()


def test_strong_scaling_shards_reassemble_the_single_gpu_result():
    """world_size 2: the two slices of the strong-scaling split, evaluated independently with a replicated basis,
    concatenate to exactly what one rank computes on the whole stream; value = all points / max time."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "bsplinekit.jl_b200"))
    sys.path.insert(0, root)
    from bsplinekit_b200 import synth
    from oracle import oracle as O
    world, points = 2, 20008
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_strong_worker, args=(r, world, port, points, q)) for r in range(world)]
    for p in procs:
        p.start()
    ally, total, value = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    k = 4
    t = O.augment_knots(synth.breakpoints(3, 50), k)
    c = 2 * synth.u_numpy(4, 0, t.size - k) - 1
    assert total == 20008 and ally.size == total
    assert np.array_equal(ally, O.spline_eval(0, k, t, c, synth.u_numpy(5, 0, total)))
    assert value == total / 2e-3 / 1e9


def test_reference_arm_runs_on_rank_0_only():
    """bench.py --impl reference under a 2-rank launch: rank 0 prints ONE JSON line (impl = reference, kind = port,
    nothing of the product loaded), every other rank exits 0 without work."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
           "--warmup", "1", "--points", "200000"]
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    r1 = subprocess.run(cmd, env=dict(env, RANK="1", LOCAL_RANK="1"), capture_output=True, text=True, timeout=300)
    assert r1.returncode == 0 and r1.stdout.strip() == ""
    r0 = subprocess.run(cmd, env=dict(env, RANK="0", LOCAL_RANK="0"), capture_output=True, text=True, timeout=300)
    assert r0.returncode == 0, r0.stderr[-2000:]
    lines = [l for l in r0.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["n_gpus"] == 2
    assert line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0 and line["scaling"] == "strong"
    assert line["config"]["points_per_step"] == 200000
