"""Time (CUDA events) one convolution shape on the tcgen05 kernel; meant to be wrapped in `ncu --set full -k regex:conv_mma`.

    python tools/prof_conv.py [n H W Cin Cout ks] [--iters 10] [--res]
"""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vistracker_b200 import _lib, ops  # noqa: E402
from vistracker_b200.weights import pack_conv  # noqa: E402

use_res = "--res" in sys.argv
args = [a for a in sys.argv[1:] if not a.startswith("--")]
if "--iters" in sys.argv:
    args.remove(sys.argv[sys.argv.index("--iters") + 1])
n, H, W, cin, cout, ks = (int(a) for a in args) if len(args) == 6 else (8, 128, 128, 256, 128, 3)
iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 10
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(n, H, W, cin, device=dev, generator=g)
w = torch.randn(cout, cin, ks, ks, device=dev, generator=g) * 0.05
pk = pack_conv(w)
planes, _ = ops.prep_split(x, None, None, True, ks // 2)
out = torch.empty(n, H, W, cout, device=dev)
stats = ops.new_stats(n, cout, dev)
res = torch.randn(n, H, W, cout, device=dev, generator=g) if use_res else None
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
P, S = _lib.ptr, _lib.stream_ptr


def run():
    _lib.call("vt_conv_mma", P(planes[0]), P(planes[1]), n, H, W, pk["cin_pad"], ks // 2, ks, P(pk["hi"]), P(pk["lo"]), cout, None, P(res), cout if use_res else 0,
              P(out), cout, P(stats), cout, S())


for _ in range(3):
    run()
times = []
for _ in range(iters):
    flush.zero_()                       # evict L2 between timed launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
times.sort()
t = times[len(times) // 2]
flops = 2.0 * n * H * W * cout * cin * ks * ks
print(f"conv_mma n={n} {H}x{W} {cin}->{cout} k{ks}: median {t*1e3:.1f} us  algorithmic {flops/t/1e9:.1f} TFLOP/s  executed-MMA {3*flops/t/1e9:.1f} TFLOP/s")
