"""Per-item role timeline of the persistent tcgen05 conv kernel (debug hook fs2_debug_set_tc_trace).

The stamps are compiled in only with FS2_TC_TRACE=1 (python fastspeech2_b200/build.py --force with that variable set); the
ablation variants of earlier rounds (tc_variant bits 1/2/4) no longer exist, so every case runs the shipped kernel."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastspeech2_b200 import _lib, ops, packing
lib = _lib.lib()
lib.fs2_debug_set_tc_trace.argtypes = [ctypes.c_void_p]
g = torch.Generator().manual_seed(0)
cases = [(128, 3, 1, 65536, False, 0), (128, 3, 1, 65536, True, 0), (128, 11, 5, 65536, False, 0),
         (32, 3, 1, 262144, False, 0), (32, 11, 1, 262144, True, 0), (256, 7, 1, 8192, False, 0)]
for (C, k, dil, T, res, variant) in cases:
    B = 16
    x = torch.randn(B, T, C, generator=g).cuda()
    w = torch.randn(k, C, C, generator=g) * (k * C) ** -0.5
    wtc = packing.pack_conv_tc(w).cuda(); w = w.cuda()
    r = torch.randn(B, T, C, generator=g).cuda() if res else None
    kw = dict(dilation=dil, pad_left=(k * dil - dil) // 2, w_tc=wtc, backend=2, res=r, tc_variant=variant)
    ops.conv1d(x, w, None, **kw); torch.cuda.synchronize()
    trace = torch.zeros(148 * 16 * 8, dtype=torch.int64, device="cuda")
    lib.fs2_debug_set_tc_trace(trace.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.conv1d(x, w, None, **kw); e1.record(); torch.cuda.synchronize()
    lib.fs2_debug_set_tc_trace(None)
    tr = trace.cpu().reshape(148, 16, 8).double()
    print(f"C={C} k={k} dil={dil} res={res} variant={variant}: kernel {e0.elapsed_time(e1)*1e3:.0f} us")
    t0 = tr[:, 0, 0].clone()
    cta = tr[0]            # CTA 0 timeline relative to its first stamp
    base = cta[0, 0]
    for il in range(3, 5):
        row = (cta[il] - base) / 1e3
        print(f"   item {il}: xform {row[0]:7.1f}->{row[1]:7.1f} | mma {row[2]:7.1f}->{row[3]:7.1f} | epi {row[4]:7.1f}->{row[5]:7.1f}   (us)")
    # steady-state averages over CTAs, items 4..12
    d = lambda a, b: (tr[:, 4:12, a] - tr[:, 4:12, b]).mean() / 1e3
    per_item = (tr[:, 12, 5] - tr[:, 4, 5]).mean() / 8e3
    print(f"   steady state: item period {per_item:.2f} us | xform span {d(1,0):.2f} | mma span {d(3,2):.2f} | epi span {d(5,4):.2f} | mma-issued -> epi start {d(4,3):.2f}")
