"""Host-side profile (cProfile) of one C4 batch through fit_recon_batch after a warm-up batch: where the Python / synchronisation time goes
between the kernels.    python tools/prof_c4_host.py [frames=96]"""
import cProfile
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import _inputs  # noqa: E402
sys.modules["tools_inputs"] = _inputs
import bench  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 96
D = bench.Dist()
c4 = bench.C4(D, frames=frames)
c4.step(c4.devd)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
c4.step(c4.devd)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
