"""filter_batch on 96 frames at different encoder chunk sizes: smaller chunks keep a layer's activations (16.8 MB per frame at 128 x 128 x 256)
inside the 126 MB L2 between the producing convolution, prep_split and the consuming convolution.

    python tools/prof_filter_chunk.py [frames]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vistracker_b200 import CHORETriplaneVisibility, default_options, resolve_dims  # noqa: E402
from vistracker_b200.recon_driver import filter_batch  # noqa: E402
from vistracker_b200.synth import synthetic_recon_batch, synthetic_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
dev = torch.device("cuda", 0)
net = CHORETriplaneVisibility(default_options(), device=dev).eval()
net.load_state_dict(synthetic_state_dict(resolve_dims(default_options()), seed=0))
net.defer_checks = True
images = synthetic_recon_batch(B, seed=4)["images"].to(dev)
ref = None
for chunk in (16, 8, 4, 2, 1):
    with torch.no_grad():
        for _ in range(2):
            filter_batch(net, images, chunk=chunk)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            filter_batch(net, images, chunk=chunk)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    maps = [m.clone() for m in net._maps]
    same = True if ref is None else all(torch.equal(a, b) for a, b in zip(maps, ref))
    if ref is None:
        ref = maps
    print(f"chunk {chunk:2d}: {ms:7.1f} ms for {B} frames = {ms / B:.2f} ms / frame; maps identical to chunk 16: {same}", flush=True)
