"""Where the generator stage of a C4 batch goes: wall clock of generate_all on 96 frames against the device-busy time of its kernels (torch profiler).

    python tools/prof_generator.py [frames] > gpurun_out/prof_generator.txt
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vistracker_b200 import CHORETriplaneVisibility, default_options, resolve_dims  # noqa: E402
from vistracker_b200.generator import GeneratorTriplaneVis  # noqa: E402
from vistracker_b200.recon_driver import generate_all  # noqa: E402
from vistracker_b200.synth import synthetic_recon_batch, synthetic_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
dev = torch.device("cuda", 0)
net = CHORETriplaneVisibility(default_options(), device=dev).eval()
net.load_state_dict(synthetic_state_dict(resolve_dims(default_options()), seed=0))
net.defer_checks = True
gen = GeneratorTriplaneVis(net, filter_val=10.0, device=dev)
h = synthetic_recon_batch(B, seed=4)
data = {k: h[k] for k in ("images", "crop_center", "body_center")}
data["images"] = data["images"].pin_memory()
for _ in range(2):
    torch.manual_seed(0)
    generate_all(gen, data, keep_maps=True)
torch.cuda.synchronize()
torch.manual_seed(0)
t0 = time.perf_counter()
generate_all(gen, data, keep_maps=True)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
print(f"generate_all on {B} frames: {wall * 1e3:.1f} ms wall")
from torch.profiler import ProfilerActivity, profile  # noqa: E402
torch.manual_seed(0)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    generate_all(gen, data, keep_maps=True)
    torch.cuda.synchronize()
ev = prof.key_averages()
rows = sorted(((e.device_time_total, e.count, e.key) for e in ev if e.device_time_total > 0 and not e.key.startswith("aten::") and "cuda" not in e.key.lower()[:4]), reverse=True)
tot = sum(r[0] for r in rows)
print(f"device-busy (kernels + memcpy): {tot / 1e3:.1f} ms")
for t, n, k in rows[:25]:
    print(f"{t / 1e3:9.2f} ms  x{n:5d}  {k[:110]}")
