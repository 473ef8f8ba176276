"""fs2_resstack per-pair mode against the two per-layer launches it replaces (64-channel stage shapes)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastspeech2_b200 import ops, packing
DEV = "cuda"
g = torch.Generator().manual_seed(0)
one = len(sys.argv) > 1
B = 16
def run(C, N, k, d, n=10):
    x = torch.randn(B, N, C, generator=g).to(DEV)
    wa = torch.randn(k, C, C, generator=g) * 0.6 * (C * k) ** -0.5; wb = torch.randn(k, C, C, generator=g) * 0.6 * (C * k) ** -0.5
    ba, bb = (torch.randn(C, generator=g) * 0.05).to(DEV), (torch.randn(C, generator=g) * 0.05).to(DEV)
    ta, tb = packing.pack_conv_tc(wa, f8=True).to(DEV), packing.pack_conv_tc(wb, f8=True).to(DEV)
    wad, wbd = wa.to(DEV), wb.to(DEV)
    y, t = torch.empty_like(x), torch.empty_like(x)
    fused = lambda: ops.resstack(x, (k,), ((d,),), [[ta]], [[ba]], [[tb]], [[bb]], alpha=1.0, out=y)
    def layered():
        ops.conv1d(x, wad, ba, dilation=d, pad_left=(k - 1) * d // 2, in_act=3, in_slope=0.1, out_act=3, out_slope=0.1, out=t, w_tc=ta, backend=2, tc_variant=1)
        ops.conv1d(t, wbd, bb, pad_left=(k - 1) // 2, res=x, out=y, w_tc=tb, backend=2, tc_variant=1)
    def timeit(fn):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    if one:
        fused(); fused(); torch.cuda.synchronize(); return
    print(f"C={C} k={k} d={d}: fused pair {timeit(fused):7.1f} us   two launches {timeit(layered):7.1f} us", flush=True)
if one:
    run(64, 129536, 3, 1)
else:
    for k in (3, 7, 11):
        for d in (1, 5):
            run(64, 129536, k, d)
    run(32, 259072, 3, 1); run(32, 259072, 7, 3)
