"""Time fs2_resstack (fused ResBlock group) against the 18 per-layer fs2_conv1d launches it replaces, per HiFi-GAN stage shape."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastspeech2_b200 import ops, packing, _lib as L
DEV = "cuda"
only = sys.argv[1] if len(sys.argv) > 1 else "all"
g = torch.Generator().manual_seed(0)
K, D = (3, 7, 11), ((1, 3, 5),) * 3
for C, N in ((32, 259072), (64, 129536)):
    B = 16
    x = torch.randn(B, N, C, generator=g).to(DEV)
    w = {}
    for f8 in (True, False):
        w[f8] = ([], [], [], [], [], [])
    raw = []
    for j, k in enumerate(K):
        for lst in w.values():
            for l in lst: l.append([])
        raw.append([])
        for d in D[j]:
            wa = torch.randn(k, C, C, generator=g) * 0.6 * (C * k) ** -0.5
            wb = torch.randn(k, C, C, generator=g) * 0.6 * (C * k) ** -0.5
            ba, bb = torch.randn(C, generator=g) * 0.05, torch.randn(C, generator=g) * 0.05
            for f8 in (True, False):
                w1, b1, w2, b2, r1, r2 = w[f8]
                w1[j].append(packing.pack_conv_tc(wa, f8=f8).to(DEV)); b1[j].append(ba.to(DEV))
                w2[j].append(packing.pack_conv_tc(wb, f8=f8).to(DEV)); b2[j].append(bb.to(DEV))
                r1[j].append(wa.to(DEV)); r2[j].append(wb.to(DEV))

    def fused():
        w1, b1, w2, b2, _, _ = w[True]
        return ops.resstack(x, K, D, w1, b1, w2, b2)

    bt, r1b, r2b, y = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)

    def layered(f8):
        w1, b1, w2, b2, ra, rb = w[f8]
        v = 1 if f8 else 0
        for j, k in enumerate(K):
            r = x
            for di, d in enumerate(D[j]):
                ops.conv1d(r, ra[j][di], b1[j][di], dilation=d, pad_left=(k - 1) * d // 2, in_act=3, in_slope=0.1, out_act=3, out_slope=0.1,
                           out=bt, w_tc=w1[j][di], backend=2, tc_variant=v)
                last = di == len(D[j]) - 1
                dst = y if last else (r2b if r is r1b else r1b)
                ops.conv1d(bt, rb[j][di], b2[j][di], pad_left=(k - 1) // 2, res=r, alpha=1 / 3 if last else 1.0, out=dst,
                           accumulate=last and j > 0, w_tc=w2[j][di], backend=2, tc_variant=v)
                r = dst
        return y

    def timeit(fn, n=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    flops = 2 * 2 * B * N * C * C * sum(K) * 3
    if only in ("all", "fused"):
        t = timeit(fused); print(f"C={C} N={N}: fused            {t:7.3f} ms  {flops / t / 1e9:7.1f} TFLOP/s", flush=True)
    if only == "all":
        a = fused().clone(); b = layered(True).clone(); torch.cuda.synchronize()
        print("   fused vs layered(f8) max abs diff", (a - b).abs().max().item())
        t = timeit(lambda: layered(True)); print(f"C={C} N={N}: 18 launches f8   {t:7.3f} ms  {flops / t / 1e9:7.1f} TFLOP/s", flush=True)
        t = timeit(lambda: layered(False)); print(f"C={C} N={N}: 18 launches s3   {t:7.3f} ms  {flops / t / 1e9:7.1f} TFLOP/s", flush=True)
