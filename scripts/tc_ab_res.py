"""A/B (same process): epilogue residual loads pipelined (variant 0) vs late (variant 8)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastspeech2_b200 import ops, packing
g = torch.Generator().manual_seed(0)
for name, C, k, T in (("s1 C128 k3 res", 128, 3, 65536), ("s1 C128 k7 res", 128, 7, 65536), ("s3 C32 k3 res", 32, 3, 262144), ("s2 C64 k11 res", 64, 11, 131072), ("s0 C256 k3 res", 256, 3, 8192)):
    x = torch.randn(16, T, C, generator=g).cuda(); r = torch.randn(16, T, C, generator=g).cuda()
    w = torch.randn(k, C, C, generator=g) * (k * C) ** -0.5
    wtc = packing.pack_conv_tc(w).cuda(); w = w.cuda()
    line = f"{name:18s}"
    for variant in (0, 8, 0, 8):
        kw = dict(pad_left=(k - 1) // 2, w_tc=wtc, backend=2, res=r, tc_variant=variant)
        ops.conv1d(x, w, None, **kw); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.conv1d(x, w, None, **kw); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        line += f" | v{variant}: {min(ts):6.0f} us"
    print(line, flush=True)
