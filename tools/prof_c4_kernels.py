"""Device time per kernel of ONE C4 batch (bench.C4.step: fit_recon_batch on 96 frames) from the CUPTI activity records of torch.profiler -- the
step runs at full speed (an `ncu --metrics gpu__time_duration.sum` pass over the same 22 k launches costs ~10 GPU-minutes, serialised and
cold-cache); kernels replayed from CUDA graphs are recorded individually.    python tools/prof_c4_kernels.py [frames=96]"""
import collections
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import _inputs  # noqa: E402
sys.modules["tools_inputs"] = _inputs
import bench  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 96
D = bench.Dist()
c4 = bench.C4(D, frames=frames)
for _ in range(2):
    c4.step(c4.devd)
torch.cuda.synchronize()
t0 = time.perf_counter()
c4.step(c4.devd)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    c4.step(c4.devd)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for e in prof.events():
    if e.device_time_total > 0:
        k = e.name.split("(")[0]
        agg[k][0] += e.device_time_total
        agg[k][1] += 1
tot = sum(v[0] for v in agg.values())
print(f"# one C4 batch of {frames} frames: {wall * 1e3:.1f} ms wall (unprofiled run), {tot / 1e3:.1f} ms of kernel + copy time, {sum(v[1] for v in agg.values())} device activities")
print(f"{'kernel':72s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}")
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{k[:72]:72s} {n:8d} {t / 1e3:10.2f} {100 * t / tot:6.1f}% {t / n:9.1f}")
