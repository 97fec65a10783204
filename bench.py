#!/usr/bin/env python
"""bench.py — headline benchmark of the spline-evaluation hot path (BASELINE.json configs[1]):
k = 4 spline on 1,000 non-uniform breakpoints, value + first derivative at 2^28 random Float64
points per GPU, one process per GPU.  One JSON line on stdout (rank 0).

  python bench.py --gpus N --steps K --warmup W             # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   # CPU restatement of the reference

A "step" is one fused pass (S(x) and S'(x)) over the 2^28 device-resident points.  `value`
is whole-job Gpoints/s with inputs resident in HBM; `e2e` is the same metric through the
host-buffer C-ABI call (pinned host arrays, H2D and D2H inside the timed region).
The batched-interpolation solve (configs[2]) is timed in the same run and reported under
"extra" with its own roofline fraction.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "bsplinekit.jl_b200"))

METRIC = "Gpoints/s spline eval (k=4, 2^28 pts, value + 1st derivative, Float64)"
UNIT = "Gpoints/s"
BYTES_PER_POINT = 24.0          # 8 B read + 2 x 8 B written
SOLVE_BYTES_PER_ELEM = 16.0     # read y + write c


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._halt.set()
        self.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def workload_c2(bk, synth):
    import numpy as np
    k = 4
    br = synth.breakpoints(3, 1000)
    B = bk.BSplineBasis(k, br)
    c = 2 * synth.u_numpy(4, 0, len(B)) - 1
    return k, B, c


def omp_threads(nthreads):
    """Use all host cores for the CPU arm (torchrun exports OMP_NUM_THREADS=1); must run before
    the oracle library is first loaded.  Returns the thread count OpenMP will actually use."""
    os.environ["OMP_NUM_THREADS"] = str(nthreads)
    try:
        import ctypes
        g = ctypes.CDLL("libgomp.so.1")
        g.omp_set_num_threads(int(nthreads))
        return int(g.omp_get_max_threads())
    except Exception:
        return nthreads


def bind_to_gpu_numa_node(index):
    """Run this rank on the CPUs next to its GPU (NVML's ideal affinity) so that its pinned host buffers
    are allocated on the NUMA node of the GPU's PCIe root — matters for the host-buffer (e2e) leg when
    several ranks stream over PCIe at once.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1 and 64 * w + b < ncpu}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:
        pass
    return None


_CPU_POINTS = {}


def cpu_reference_rate(npts_sample, nthreads):
    """The reference algorithm on the host cores: de Boor value + derivative spline, OpenMP over
    points (oracle/bsk_oracle.c).  Returns (Gpoints/s, seconds)."""
    import numpy as np
    from oracle import oracle as O
    from bsplinekit_b200 import synth
    k = 4
    br = synth.breakpoints(3, 1000)
    t = O.augment_knots(br, k)
    c = 2 * synth.u_numpy(4, 0, t.size - k) - 1
    kd, td, cd = O.spline_derivative(0, k, t, c, 1)
    if npts_sample not in _CPU_POINTS:          # the sample is generated once, outside the timed region
        _CPU_POINTS.clear()
        _CPU_POINTS[npts_sample] = synth.u_numpy(5, 0, npts_sample)
    x = _CPU_POINTS[npts_sample]
    O.spline_eval(0, k, t, c, x[:1000])
    t0 = time.perf_counter()
    O.spline_eval(0, k, t, c, x)
    O.spline_eval(0, kd, td, cd, x)
    dt = time.perf_counter() - t0
    return npts_sample / dt / 1e9, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = omp_threads(os.cpu_count() or 1)
    sample = 1 << 26
    rates, times = [], []
    for _ in range(args.warmup):
        cpu_reference_rate(1 << 20, cores)
    for _ in range(args.steps):
        r, dt = cpu_reference_rate(sample, cores)
        rates.append(r)
        times.append(dt)
    tot = sum(times)
    value = sample * len(times) / tot / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "C2 cubic eval value+derivative, k=4, 1000 non-uniform breakpoints",
                   "points_per_step": sample, "note": "bounded sample of the 2^28-point workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "2^26 points per step, OpenMP over points, all host cores"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import numpy as np
    import torch
    import bsplinekit_b200 as bk
    from bsplinekit_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    bind_to_gpu_numa_node(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    npts = args.points
    k, B, c = workload_c2(bk, synth)
    S = bk.Spline(B, c)
    # each rank generates its own slice of the point stream on its device
    xd = synth.u_torch(5, rank * npts, npts)
    v = torch.empty_like(xd)
    dv = torch.empty_like(xd)
    outs = [v, dv]
    algo = {"auto": bk.EVAL_AUTO, "pp": bk.EVAL_PP, "deboor": bk.EVAL_DEBOOR}[args.algo]

    for _ in range(args.warmup):
        S.evaluate(xd, (0, 1), algo=algo, out=outs)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        S.evaluate(xd, (0, 1), algo=algo, out=outs)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    ms_total = max_over_ranks(ms_total)
    ms_step = ms_total / args.steps
    value = world * npts / (ms_step * 1e-3) / 1e9
    launches_per_step = 1 if args.algo != "deboor" else 2
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_POINT * npts / (ms_step * 1e-3) / 1e9      # per GPU, GB/s
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get("k_spline_cell_bytes_per_launch")
            if t is not None and args.algo != "deboor":
                traffic = t * (npts / float(1 << 28))      # captured on 2^28 points
    except Exception:
        pass

    # the same access mix without the table lookups (n doubles in, 2 n out, same launch shape): what HBM
    # gives a 1 : 2 read : write stream on this GPU, measured the same way beside the kernel
    stream_mix = None
    try:
        import ctypes
        from bsplinekit_b200 import _lib as _bl
        sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        w0 = torch.empty_like(v); w1 = torch.empty_like(dv)      # scratch: the results stay intact
        probe = lambda: _bl.check(_bl.lib.bsk_probe_stream_1r2w(ctypes.c_void_p(xd.data_ptr()), ctypes.c_void_p(w0.data_ptr()),
                                                                ctypes.c_void_p(w1.data_ptr()), npts, sp))
        for _ in range(3):
            probe()
        p0 = torch.cuda.Event(enable_timing=True)
        p1 = torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(10):
            probe()
        p1.record()
        torch.cuda.synchronize()
        stream_mix = BYTES_PER_POINT * npts / (p0.elapsed_time(p1) / 10 * 1e-3) / 1e9
        del w0, w1
    except Exception as exc:                 # the probe is informative only
        print("stream probe skipped: %r" % (exc,), file=sys.stderr)

    # parity spot check on what was just computed (not timed)
    chk = {}
    if rank == 0:
        from oracle import oracle as O
        m = 1 << 16
        xh = xd[:m].cpu().numpy()
        ref = O.spline_eval(0, k, B.knots(), c, xh)
        kd, td, cd = O.spline_derivative(0, k, B.knots(), c, 1)
        refd = O.spline_eval(0, kd, td, cd, xh)
        chk = {"value_relerr": float(np.max(np.abs(v[:m].cpu().numpy() - ref)) / np.max(np.abs(ref))),
               "deriv_relerr": float(np.max(np.abs(dv[:m].cpu().numpy() - refd)) / np.max(np.abs(refd)))}

    # ---- e2e: host buffers through the C-ABI host call (H2D + kernel + D2H, pipelined) ----------
    e2e = None
    if not args.no_e2e:
        ne = min(npts, args.e2e_points)
        xh = torch.empty(ne, dtype=torch.float64, pin_memory=True)
        xh.copy_(xd[:ne])
        oh = [torch.empty(ne, dtype=torch.float64, pin_memory=True) for _ in range(2)]
        xnp, onp = xh.numpy(), [o.numpy() for o in oh]
        S.evaluate(xnp, (0, 1), algo=algo, out=onp)       # warm-up
        barrier()
        t0 = time.perf_counter()
        ke = max(1, min(args.steps, args.e2e_steps))
        for _ in range(ke):
            S.evaluate(xnp, (0, 1), algo=algo, out=onp)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0) / ke
        assert np.array_equal(onp[0][:1024], v[:1024].cpu().numpy())
        e2e = {"value": world * ne / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": 8 * ne,
               "d2h_bytes_per_step": 16 * ne, "points_per_step": ne, "steps": ke,
               "ms_per_step": 1e3 * dt}
        del xh, oh

    # ---- extra: batched interpolation solve, configs[2] -------------------------------------------
    extra = {}
    if not args.no_solve:
        del xd, v, dv
        torch.cuda.empty_cache()
        n, kk, nrhs = 4096, 6, args.nrhs
        xdata = synth.breakpoints(6, n)
        Bi = bk.BSplineBasis(kk, bk.make_knots(xdata, kk), augment=False)
        t_f0 = time.perf_counter()
        itp = bk.SplineInterpolation(Bi, xdata)          # assemble + LU, once
        torch.cuda.synchronize()
        factor_first_ms = 1e3 * (time.perf_counter() - t_f0)   # first call in the process: + lazy kernel loading
        del itp
        t_f0 = time.perf_counter()
        itp = bk.SplineInterpolation(Bi, xdata)
        torch.cuda.synchronize()
        factor_ms = 1e3 * (time.perf_counter() - t_f0)
        Y0 = (2 * synth.u_torch(7, rank * n * nrhs, n * nrhs) - 1).reshape(n, nrhs)
        Y = torch.empty_like(Y0)
        res = {}
        for name, salgo in (("windowed", bk.SOLVE_WINDOWED), ("twopass", bk.SOLVE_TWOPASS)):
            if name == "windowed" and itp.C.halo == 0:
                continue
            tot = 0.0
            for it in range(args.warmup + args.solve_steps):
                Y.copy_(Y0)
                barrier()
                s0 = torch.cuda.Event(enable_timing=True)
                s1 = torch.cuda.Event(enable_timing=True)
                s0.record()
                itp.C.ldiv_(Y, algo=salgo)
                s1.record()
                torch.cuda.synchronize()
                if it >= args.warmup:
                    tot += s0.elapsed_time(s1)
            ms = max_over_ranks(tot / args.solve_steps)
            res[name] = {"ms": ms, "rhs_per_s": world * nrhs / (ms * 1e-3),
                         "achieved_gbs": SOLVE_BYTES_PER_ELEM * n * nrhs / (ms * 1e-3) / 1e9,
                         "frac": SOLVE_BYTES_PER_ELEM * n * nrhs / (ms * 1e-3) / 1e9 / peak}
            if rank == 0 and name == "windowed":
                keep = Y[:, :256].cpu().numpy()
        if rank == 0:
            from oracle import oracle as O
            band, _ = O.collocation_matrix(0, kk, Bi.knots(), xdata)
            O.banded_lu(band)
            ref = O.banded_solve(band, Y0[:, :256].cpu().numpy().copy())
            got = keep if "windowed" in res else Y[:, :256].cpu().numpy()
            chk["solve_relerr"] = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
        best = min(res.values(), key=lambda r: r["ms"])
        extra = {"c3_batched_solve": {"metric": "RHS/s (n=4096, k=6, %d RHS per GPU, Float64)" % nrhs,
                                      "value": best["rhs_per_s"], "unit": "RHS/s",
                                      "roofline_frac": best["frac"], "algos": res,
                                      "factor_once_ms": factor_ms, "factor_first_call_ms": factor_first_ms,
                                      "halo_rows": itp.C.halo}}
        # du = F \ (L u), mul! + ldiv! of a collocation time step in one pass
        Lm = bk.collocation_matrix(Bi, xdata, bk.Derivative(2))
        tot = 0.0
        for it in range(args.warmup + args.solve_steps):
            barrier()
            s0 = torch.cuda.Event(enable_timing=True)
            s1 = torch.cuda.Event(enable_timing=True)
            s0.record()
            itp.C.ldiv_mul(Lm, Y0, out=Y)
            s1.record()
            torch.cuda.synchronize()
            if it >= args.warmup:
                tot += s0.elapsed_time(s1)
        ms = max_over_ranks(tot / args.solve_steps)
        if rank == 0:
            yref = O.banded_matvec(O.collocation_matrix(0, kk, Bi.knots(), xdata, 2)[0], Y0[:, :256].cpu().numpy())
            O.banded_solve(band, yref)
            chk["fused_matvec_solve_relerr"] = float(np.max(np.abs(Y[:, :256].cpu().numpy() - yref)) / np.max(np.abs(yref)))
        extra["n1_fused_matvec_solve"] = {"metric": "RHS/s, du = F \\ (D2 u) (n=4096, k=6, %d columns per GPU)" % nrhs,
                                          "ms": ms, "value": world * nrhs / (ms * 1e-3), "unit": "RHS/s",
                                          "roofline_frac": SOLVE_BYTES_PER_ELEM * n * nrhs / (ms * 1e-3) / 1e9 / peak}
        del Lm
        # all `nrhs` interpolants (coefficient block left by the solve) at 4096 points
        SBn = bk.SplineBatch(Bi, Y)
        xsn = synth.u_torch(5, 0, 4096)
        outn = None
        tot = 0.0
        for it in range(args.warmup + args.solve_steps):
            barrier()
            s0 = torch.cuda.Event(enable_timing=True)
            s1 = torch.cuda.Event(enable_timing=True)
            s0.record()
            outn = SBn(xsn)
            s1.record()
            torch.cuda.synchronize()
            if it >= args.warmup:
                tot += s0.elapsed_time(s1)
        ms = max_over_ranks(tot / args.solve_steps)
        if rank == 0:
            cn = Y[:, 7].cpu().numpy().copy()
            refn = O.spline_eval(0, kk, Bi.knots(), cn, xsn.cpu().numpy())
            chk["vector_eval_relerr"] = float(np.max(np.abs(outn[:, 7].cpu().numpy() - refn)) / np.max(np.abs(refn)))
        nb_ = 8.0 * 4096 * nrhs + 8.0 * n * nrhs                # values written + coefficient block read once
        extra["n2_vector_valued_eval"] = {"metric": "spline values/s, %d splines sharing one basis x 4096 points" % nrhs,
                                          "ms": ms, "value": world * 4096.0 * nrhs / (ms * 1e-3), "unit": "values/s",
                                          "roofline_frac": nb_ / (ms * 1e-3) / 1e9 / peak}
        del SBn, outn

    # ---- extra: configs[3] (periodic cubic) and configs[4] (Robin k = 8 + Float32 sweep) ------------
    if not args.no_extra_configs:
        def timed(fn, reps, setup=None):
            tot = 0.0
            for it in range(reps + 1):
                if setup is not None:
                    setup()
                torch.cuda.synchronize()
                a = torch.cuda.Event(enable_timing=True)
                b_ = torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b_.record()
                torch.cuda.synchronize()
                if it > 0:
                    tot += a.elapsed_time(b_)
            return max_over_ranks(tot / reps)

        torch.cuda.empty_cache()
        n4 = 10000
        br4 = synth.breakpoints(8, n4 + 1)
        B4 = bk.PeriodicBSplineBasis(4, br4)
        M4 = bk.collocation_matrix(B4, bk.collocation_points(B4))
        nrhs4 = 16384
        D0 = (2 * synth.u_torch(9, rank * n4 * nrhs4, n4 * nrhs4) - 1).reshape(n4, nrhs4)
        D = torch.empty_like(D0)
        ms = timed(lambda: M4.ldiv_(D), 5, setup=lambda: D.copy_(D0))
        extra["c4_cyclic_solve"] = {"metric": "RHS/s (periodic cubic, n=10^4, 16384 RHS per GPU)", "ms": ms,
                                    "value": world * nrhs4 / (ms * 1e-3), "unit": "RHS/s",
                                    "roofline_frac": 16.0 * n4 * nrhs4 / (ms * 1e-3) / 1e9 / peak}
        S4 = bk.Spline(B4, D[:, 0].cpu().numpy())
        del D, D0
        x4 = synth.u_torch(9, (1 << 40) + rank * npts, npts)
        o4 = [torch.empty_like(x4)]
        ms = timed(lambda: S4.evaluate(x4, (0,), out=o4), 5)
        extra["c4_periodic_eval"] = {"metric": "Gpoints/s (periodic cubic, 10^4 knots, 2^28 F64 points)", "ms": ms,
                                     "value": world * npts / (ms * 1e-3) / 1e9, "unit": "Gpoints/s",
                                     "roofline_frac": 16.0 * npts / (ms * 1e-3) / 1e9 / peak}
        del x4, o4
        B5 = bk.BSplineBasis(8, synth.breakpoints(10, 99996))
        R5 = bk.RecombinedBSplineBasis(B5, bk.Derivative(0) + 3 * bk.Derivative(1))
        x5 = torch.from_numpy(bk.collocation_points(R5)).cuda()
        ms = timed(lambda: bk.collocation_matrix(R5, x5, bk.Derivative(2)), 5)
        extra["c5_collocation_assembly"] = {"metric": "ms per collocation_matrix(R, x, Derivative(2)), 10^5 x k=8 "
                                                      "(assembly + band handle; launch-bound)", "ms": ms}
        B5f = bk.BSplineBasis(8, synth.breakpoints(10, 99996).astype(np.float32))
        S5f = bk.Spline(B5f, (2 * synth.u_numpy(4, 0, len(B5f)) - 1).astype(np.float32))
        x5f = synth.u_torch(11, rank * npts, npts, dtype=torch.float32)
        o5 = [torch.empty_like(x5f)]
        ms = timed(lambda: S5f.evaluate(x5f, (0,), out=o5), 5)
        extra["c5_float32_eval"] = {"metric": "Gpoints/s (k=8, 10^5 knots, 2^28 F32 points)", "ms": ms,
                                    "value": world * npts / (ms * 1e-3) / 1e9, "unit": "Gpoints/s",
                                    "roofline_frac": 8.0 * npts / (ms * 1e-3) / 1e9 / peak}
        del x5f, o5
        # Float32 batched interpolation (C3 shape in Float32; BANSLV order, two passes: 16 B/element moved,
        # 8 B/element algorithmic) and evaluation of an order-6 Float64 spline (shared-memory cell table
        # with the jump encoding of knot cells)
        xd32 = synth.breakpoints(6, 4096).astype(np.float32)
        B32 = bk.BSplineBasis(6, bk.make_knots(xd32, 6), augment=False)
        itp32 = bk.SplineInterpolation(B32, xd32)
        nr32 = 65536
        Y32 = (2 * synth.u_torch(7, rank * 4096 * nr32, 4096 * nr32, dtype=torch.float32) - 1).reshape(4096, nr32)
        W32 = torch.empty_like(Y32)
        ms = timed(lambda: itp32.C.ldiv_(W32), 5, setup=lambda: W32.copy_(Y32))
        extra["float32_batched_solve"] = {"metric": "RHS/s (n=4096, k=6, 65536 RHS per GPU, Float32, BANSLV order)",
                                          "ms": ms, "value": world * nr32 / (ms * 1e-3), "unit": "RHS/s",
                                          "roofline_frac": 8.0 * 4096 * nr32 / (ms * 1e-3) / 1e9 / peak}
        if rank == 0:
            from oracle import oracle as O32
            band32, _ = O32.collocation_matrix_f32(0, 6, B32.knots(), xd32)
            O32.banded_lu(band32)
            ref32 = O32.banded_solve(band32, Y32[:, :128].cpu().numpy().copy())
            chk["float32_solve_bitexact"] = bool(np.array_equal(W32[:, :128].cpu().numpy(), ref32))
        del Y32, W32, itp32
        B6 = bk.BSplineBasis(6, synth.breakpoints(3, 1000))
        S6 = bk.Spline(B6, 2 * synth.u_numpy(4, 0, len(B6)) - 1)
        x6 = synth.u_torch(5, rank * npts, npts)
        o6 = [torch.empty_like(x6), torch.empty_like(x6)]
        ms = timed(lambda: S6.evaluate(x6, (0, 1), out=o6), 5)
        extra["order6_eval"] = {"metric": "Gpoints/s (k=6, 1000 breakpoints, value + 1st derivative, 2^28 F64 points)",
                                "ms": ms, "value": world * npts / (ms * 1e-3) / 1e9, "unit": "Gpoints/s",
                                "roofline_frac": 24.0 * npts / (ms * 1e-3) / 1e9 / peak}
        if rank == 0:
            from oracle import oracle as O6
            x6h = x6[:1 << 16].cpu().numpy()
            c6 = 2 * synth.u_numpy(4, 0, len(B6)) - 1
            r6 = O6.spline_eval(0, 6, B6.knots(), c6, x6h)
            chk["order6_value_relerr"] = float(np.max(np.abs(o6[0][:1 << 16].cpu().numpy() - r6)) / np.max(np.abs(r6)))
        del x6, o6

    # ---- extra: configs[0] (the reference's own CPU-runnable case), in full, host arrays in and out --
    if not args.no_extra_configs and rank == 0:
        xdata1 = (np.arange(11.0)) ** 2
        ydata1 = synth.u_numpy(1, 0, 11)
        xs1 = 100.0 * synth.u_numpy(2, 0, 1000000)
        t0 = time.perf_counter()
        itp1 = bk.interpolate(xdata1, ydata1, bk.BSplineOrder(4))
        v1 = itp1(xs1)                                   # first call builds the tables
        torch.cuda.synchronize()
        first_ms = 1e3 * (time.perf_counter() - t0)
        t0 = time.perf_counter()
        for _ in range(5):
            itp1.interpolate_(ydata1)
            v1 = itp1(xs1)
        torch.cuda.synchronize()
        rep_ms = 1e3 * (time.perf_counter() - t0) / 5
        from oracle import oracle as O1
        t1k = O1.make_knots(xdata1, 4)
        band1, _ = O1.collocation_matrix(0, 4, t1k, xdata1)
        O1.banded_lu(band1)
        c1 = O1.banded_solve(band1, ydata1.copy().reshape(-1, 1)).ravel()
        omp_threads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        ref1 = O1.spline_eval(0, 4, t1k, c1, xs1)
        cpu_ms = 1e3 * (time.perf_counter() - t0)
        chk["c1_relerr"] = float(np.max(np.abs(v1 - ref1)) / np.max(np.abs(ref1)))
        extra["c1_interpolate_eval_1e6"] = {"metric": "ms, interpolate((0:10).^2, y, k=4) + S.(xs) at 10^6 host points "
                                                      "(H2D + kernel + D2H)", "first_call_ms": first_ms, "ms": rep_ms,
                                            "cpu_port_eval_ms": cpu_ms}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = omp_threads(os.cpu_count() or 1)
        sample = 1 << 27
        r, dt = cpu_reference_rate(sample, cores)
        cpu = {"value": r, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "2^27 of the 2^28 points, value + derivative, OpenMP over points (%.1f s)" % dt}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2: k=4 spline, 1000 non-uniform breakpoints, value + 1st derivative "
                                   "at %d random Float64 points per GPU" % npts,
                       "points_per_gpu": npts, "algo": args.algo,
                       "l2": "inputs (%.2f GB per step) larger than the 126 MB L2" % (8e-9 * npts),
                       "parallelism": "points sharded, basis replicated, no collective"},
            "clocks": clocks, "gpu_launches": args.steps * launches_per_step,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "stream_1r2w_gbs": stream_mix,
                         "frac_of_stream_1r2w": (achieved / stream_mix) if stream_mix else None,
                         "kernel": "k_spline_cell<4,0,double,3>" if args.algo != "deboor" else "k_spline_deboor<4,0,double>",
                         "algorithmic_bytes_per_launch": BYTES_PER_POINT * npts},
            "e2e": e2e, "cpu_baseline": cpu, "parity": chk, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=1 << 28)
    ap.add_argument("--algo", default="auto", choices=["auto", "pp", "deboor"])
    ap.add_argument("--nrhs", type=int, default=65536)
    ap.add_argument("--solve-steps", type=int, default=10)
    ap.add_argument("--e2e-points", type=int, default=1 << 28)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-solve", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
