"""Run the BASELINE config-2 query (8 frames x 10k points) a few times; wrap in `ncu -k regex:query_fwd_tc` for a capture."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vistracker_b200 import CHORETriplaneVisibility, default_options, resolve_dims  # noqa: E402
from vistracker_b200.synth import synthetic_frames, synthetic_state_dict  # noqa: E402

dev = torch.device("cuda", 0)
net = CHORETriplaneVisibility(default_options(), device=dev).eval()
net.load_state_dict(synthetic_state_dict(resolve_dims(default_options()), seed=0))
net.defer_checks = True
images, points, crop, body = synthetic_frames(8, size=512, seed=2, n_points=10000, jitter=True)
net.filter(images.to(dev))
pts, cc, bc = points.to(dev), crop.to(dev), body.to(dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); net._query_raw(pts, cc, bc); e1.record(); torch.cuda.synchronize()
    print(f"query 8 x 10000 points: {e0.elapsed_time(e1):.3f} ms")
if "--bwd" in sys.argv:
    g = torch.randn(8, 29, 10000, device=dev)
    for mask in (5, 1):
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); net._query_backward(pts, cc, bc, g, head_mask=mask); e1.record(); torch.cuda.synchronize()
            print(f"query backward 8 x 10000 points, head mask {mask}: {e0.elapsed_time(e1):.3f} ms")
