"""One fs2_conv_post launch at the generator's bench shape (16 x 259072 samples x 32 channels, 7 taps), for `ncu --set full`."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastspeech2_b200 import ops
x = torch.randn(16, 259072, 32, device="cuda"); w = torch.randn(7, 32, device="cuda") * 0.1; b = torch.zeros(1, device="cuda")
for _ in range(3):
    ops.conv_post(x, w, b, 0.01)
torch.cuda.synchronize()
