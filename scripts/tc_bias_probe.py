"""Is the tensor core's fp32 accumulation biased (round-toward-zero)?  All-positive operands, growing K."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastspeech2_b200 import ops, packing
g = torch.Generator().manual_seed(0)
for Cin in (16, 64, 256, 1024, 4096):
    x = torch.rand(1, 128, Cin, generator=g) + 0.5
    w = (torch.rand(1, Cin, 128, generator=g) + 0.5) / Cin
    want = (x.double() @ w.double()[0])
    for label, kw in (("simt", dict(backend=1)), ("tc", dict(backend=2, w_tc=packing.pack_conv_tc(w).cuda()))):
        got = ops.conv1d(x.cuda(), w.cuda(), None, **kw).cpu().double()
        rel = (got - want) / want
        print(f"Cin {Cin:5d} {label:4s} rel err mean {rel.mean():+.3e}  abs-mean {rel.abs().mean():.3e}  max {rel.abs().max():.3e}")
