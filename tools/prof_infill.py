"""HVOP-Net timing: one autoregressive in-filling of a synthetic sequence, eager launches and CUDA-graph replay.

    python tools/prof_infill.py [--short]      (--short: one eager 400-frame pass, for an ncu launch list)
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vistracker_b200.infill import CondMotionInfillAutoreg, ConditionalMInfiller, default_infill_options  # noqa: E402
from vistracker_b200.synth import synthetic_infill_sequence, synthetic_infill_state_dict  # noqa: E402

short = "--short" in sys.argv
opt = default_infill_options()
net = ConditionalMInfiller(opt, device="cuda:0").load_state_dict(synthetic_infill_state_dict(opt, 1))
for L in ((400,) if short else (400, 1500)):
    seq = [torch.from_numpy(a).cuda() for a in synthetic_infill_sequence(L, seed=3)]
    occ = seq[4].cpu().numpy()
    for graph in ((False,) if short else (False, True)):
        drv = CondMotionInfillAutoreg(net, use_graph=graph)
        n = 1 if short else 5
        for _ in range(0 if short else 2):
            drv.infill(*seq[:4], occ)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            drv.infill(*seq[:4], occ)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        clips = len(drv.clip_plan(L))
        print(f"L={L} graph={graph}: {dt * 1e3:.2f} ms / sequence, {clips} clips, {dt * 1e6 / clips:.0f} us / clip, {L / dt:.0f} frames/s")
