"""A/B (same process): MT=4 work items for narrow layers vs MT=2 (forced through the debug tuning hook)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastspeech2_b200 import _lib, ops, packing
lib = _lib.lib()
lib.fs2_debug_set_tc_tuning.argtypes = [ctypes.c_int] * 4
g = torch.Generator().manual_seed(0)
for name, Cin, N, k, T, res in (("s3 C32 k3", 32, 32, 3, 262144, False), ("s3 C32 k3 res", 32, 32, 3, 262144, True), ("s3 C32 k11", 32, 32, 11, 262144, False),
                                ("s2 C64 k3", 64, 64, 3, 131072, False), ("s2 C64 k7 res", 64, 64, 7, 131072, True), ("s2 C64 k11", 64, 64, 11, 131072, False),
                                ("ups3 C64->32 k2", 64, 32, 2, 131072, False), ("ups2 C128->64 k2", 128, 64, 2, 65536, False)):
    x = torch.randn(16, T, Cin, generator=g).cuda(); r = torch.randn(16, T, N, generator=g).cuda() if res else None
    w = torch.randn(k, Cin, N, generator=g) * (k * Cin) ** -0.5
    wtc = packing.pack_conv_tc(w).cuda(); w = w.cuda()
    line = f"{name:18s}"
    ref = None
    for force2 in (0, -1, 0, -1):
        lib.fs2_debug_set_tc_tuning(0, 0, 0, force2)
        kw = dict(pad_left=(k - 1) // 2, w_tc=wtc, backend=2, res=r)
        y = ops.conv1d(x, w, None, **kw); torch.cuda.synchronize()
        if ref is None: ref = y
        else: assert torch.equal(ref, y) or (ref - y).abs().max() < 1e-4
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.conv1d(x, w, None, **kw); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        line += f" | {'MT4' if force2 == 0 else 'MT2'}: {min(ts):6.0f} us"
    lib.fs2_debug_set_tc_tuning(0, 0, 0, 0)
    print(line, flush=True)
