"""Host issue cost vs device time per step of S.evaluate on device-resident points (4096 and 2^25 points, 20 and
200 back-to-back steps): how much of a short step is launch overhead.  Run on the GPU box."""
import sys, os, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/bsplinekit.jl_b200")
import torch, numpy as np
import bsplinekit_b200 as bk
from bsplinekit_b200 import synth
B = bk.BSplineBasis(4, synth.breakpoints(3, 1000))
S = bk.Spline(B, 2 * synth.u_numpy(4, 0, len(B)) - 1)
for npts in (1 << 12, 1 << 25):
    xd = synth.u_torch(5, 0, npts)
    outs = [torch.empty_like(xd), torch.empty_like(xd)]
    for _ in range(5): S.evaluate(xd, (0, 1), out=outs)
    torch.cuda.synchronize()
    for steps in (20, 200):
        for rep in range(3):
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(steps): S.evaluate(xd, (0, 1), out=outs)
            e1.record()
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            print("npts %d steps %d: host issue %.1f us/call, device %.2f us/step" % (npts, steps, (t1 - t0) / steps * 1e6, e0.elapsed_time(e1) / steps * 1e3), flush=True)
