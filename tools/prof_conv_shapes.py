"""Per-shape timing of every tensor-core convolution one `filter()` step launches (BASELINE config 2: B=8, 512x512).

Records the (n, H, W, Cin, Cout, ks, residual?, stats?) of each `vt_conv_mma` call of a real filter pass, then times each unique
shape alone (CUDA events, L2 flushed between launches) for the kernel variants selected by environment switches, and prints the
per-shape table plus the sum over the step.

    python tools/prof_conv_shapes.py [--batch 8] [--modes persist,plain] [--iters 7]
"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vistracker_b200 import CHORETriplaneVisibility, _lib, default_options, ops, resolve_dims  # noqa: E402
from vistracker_b200 import encoder as E  # noqa: E402
from vistracker_b200.synth import synthetic_frames, synthetic_state_dict  # noqa: E402
from vistracker_b200.weights import pack_conv  # noqa: E402


def arg(name, default):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default


B = int(arg("--batch", 8))
modes = arg("--modes", "persist,plain").split(",")
iters = int(arg("--iters", 7))
MODE_ENV = {"persist": {"VT_CONV_PERSIST": "1", "VT_CONV_STRIP": "0", "VT_CONV_PREFETCH": "0"},
            "e8": {"VT_CONV_PERSIST": "6", "VT_CONV_STRIP": "0", "VT_CONV_PREFETCH": "0"},
            "nostrip": {"VT_CONV_PERSIST": "3", "VT_CONV_STRIP": "0", "VT_CONV_PREFETCH": "0"},
            "pf2": {"VT_CONV_PERSIST": "1", "VT_CONV_STRIP": "0", "VT_CONV_PREFETCH": "2"},
            "pf4": {"VT_CONV_PERSIST": "1", "VT_CONV_STRIP": "0", "VT_CONV_PREFETCH": "4"},
            "nores": {"VT_CONV_PERSIST": "2", "VT_CONV_STRIP": "0", "VT_CONV_PREFETCH": "0"},
            "plain": {"VT_CONV_PERSIST": "0", "VT_CONV_STRIP": "0"},
            "strip": {"VT_CONV_PERSIST": "0", "VT_CONV_STRIP": "1"}, "pair": {"VT_CONV_PERSIST": "0", "VT_CONV_STRIP": "2"}}

dev = torch.device("cuda", 0)
dims = resolve_dims(default_options())
net = CHORETriplaneVisibility(default_options(), device="cuda:0").eval()
net.load_state_dict(synthetic_state_dict(dims, seed=0))
images, *_ = synthetic_frames(B, size=512, seed=0, n_points=16)

shapes = collections.Counter()
orig = E.HGEncoder._conv


def spy(self, op, name, out, bias=None, res=None, stats=None, out2=None, res2=None):
    a = op.act
    pk = self.conv[name]
    if E.mma_tileable(a.H, a.W) and not self.force_ffma:
        shapes[(a.n, a.H, a.W, pk["cin"], pk["cout"], pk["ks"], res is not None, stats is not None)] += 1
    return orig(self, op, name, out, bias, res, stats, out2, res2)


E.HGEncoder._conv = spy
net.use_graph = False
net.filter(images.to(dev))
torch.cuda.synchronize()
E.HGEncoder._conv = orig

flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
P, S = _lib.ptr, _lib.stream_ptr
g = torch.Generator(device=dev).manual_seed(0)
totals = {m: 0.0 for m in modes}
tot_flop = 0.0
print(f"# B={B}; columns: count x shape, then per mode: median us per launch (executed-MMA TFLOP/s)")
for key, cnt in sorted(shapes.items(), key=lambda kv: -kv[1] * kv[0][0] * kv[0][1] * kv[0][2] * kv[0][3] * kv[0][4] * kv[0][5] ** 2):
    n, H, W, cin, cout, ks, has_res, has_stats = key
    x = torch.randn(n, H, W, cin, device=dev, generator=g)
    w = torch.randn(cout, cin, ks, ks, device=dev, generator=g) * 0.05
    pk = pack_conv(w)
    planes, _ = ops.prep_split(x, None, None, True, ks // 2)
    out = torch.empty(n, H, W, cout, device=dev)
    res = torch.randn(n, H, W, cout, device=dev, generator=g) if has_res else None
    stats = ops.new_stats(n, cout, dev) if has_stats else None
    flops = 2.0 * n * H * W * cout * cin * ks * ks
    tot_flop += flops * cnt
    row = f"{cnt:3d} x n={n:2d} {H:3d}x{W:3d} {cin:3d}->{cout:3d} k{ks} res={int(has_res)} st={int(has_stats)}"
    for m in modes:
        os.environ.update(MODE_ENV[m])

        def run():
            _lib.call("vt_conv_mma", P(planes[0]), P(planes[1]), n, H, W, pk["cin_pad"], ks // 2, ks, P(pk["hi"]), P(pk["lo"]), cout, None,
                      P(res) if has_res else None, cout if has_res else 0, P(out), cout, P(stats) if has_stats else None,
                      cout if has_stats else 0, S())
        for _ in range(2):
            run()
        times = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        times.sort()
        t = times[len(times) // 2]
        totals[m] += t * cnt
        row += f" | {m}: {t * 1e3:7.1f} us ({3 * flops / t / 1e9:6.0f})"
    print(row)
print("# per filter step: " + ", ".join(f"{m} {totals[m]:.2f} ms ({3 * tot_flop / totals[m] / 1e9:.0f} TFLOP/s executed, "
                                         f"{tot_flop / totals[m] / 1e9:.0f} algorithmic)" for m in modes))
