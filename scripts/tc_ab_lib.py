"""A/B of two builds of libfs2b200.so inside one process (box-to-box noise is 10-20 %): the in-tree library against
a second build passed on the command line (e.g. the previous commit built into scratch_ab/, which is not tracked).  Same inputs, outputs must match bit for bit.

usage: python scripts/tc_ab_lib.py [old.so]
"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fastspeech2_b200 import _lib as L, packing

old_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "scratch_ab", "libfs2b200_old.so")
libs = {"new": C.CDLL(L.LIB_PATH), "old": C.CDLL(old_path)}
for lib in libs.values():
    lib.fs2_conv1d.restype = C.c_int
    lib.fs2_conv1d.argtypes = [C.c_void_p, C.c_void_p]
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run(lib, x, w, wtc, y, *, bias=None, k=1, dil=1, res=None, acc=False, lens=None, in_act=0, out_act=0, alpha=1.0):
    B, T, Cin = x.shape
    N = w.shape[2]
    a = L.Conv1dArgs(x=x.data_ptr(), x_batch_stride=x.stride(0), x_row_stride=x.stride(1), B=B, T=T, Cin=Cin, w=w.data_ptr(), bias=L.ptr(bias),
                     N=N, taps=k, dilation=dil, pad_left=(k - 1) * dil // 2, w_tc=wtc.data_ptr(), backend=L.CONV_TC, tc_variant=0,
                     in_act=in_act, in_slope=0.1, out_act=out_act, out_slope=0.1, res=L.ptr(res),
                     res_batch_stride=res.stride(0) if res is not None else 0, res_row_stride=res.stride(1) if res is not None else 0,
                     alpha=alpha, accumulate=int(acc), row_lens=L.ptr(lens), y=y.data_ptr(), y_batch_stride=y.stride(0), y_row_stride=y.stride(1))
    rc = lib.fs2_conv1d(C.byref(a), stream)
    assert rc == 0, rc


g = torch.Generator().manual_seed(0)
# correctness first: ragged / partial tiles / every epilogue mode, new == old bit for bit
for (B, T, Cin, N, k, dil, res, acc, in_act, out_act) in ((3, 300, 64, 96, 3, 1, True, True, 3, 0), (2, 1000, 256, 256, 5, 1, True, False, 0, 2), (5, 77, 32, 32, 7, 3, False, False, 3, 3),
                                                           (2, 515, 128, 64, 11, 5, False, True, 3, 1), (1, 129, 80, 512, 5, 1, True, True, 0, 0), (2, 700, 256, 80, 1, 1, False, False, 0, 0),
                                                           (4, 2051, 32, 32, 3, 1, True, False, 3, 0), (2, 4100, 64, 64, 7, 1, True, True, 3, 0)):
    x = (torch.randn(B, T, Cin, generator=g) * 3).cuda(); r = torch.randn(B, T, N, generator=g).cuda() if res else None
    lens = torch.randint(1, T + 1, (B,), generator=g).int().cuda()
    w = torch.randn(k, Cin, N, generator=g) * (k * Cin) ** -0.5
    b = torch.randn(N, generator=g).cuda()
    wtc = packing.pack_conv_tc(w).cuda(); w = w.cuda()
    outs = {}
    for name, lib in libs.items():
        y = torch.full((B, T, N), 0.25, device="cuda")
        run(lib, x, w, wtc, y, bias=b, k=k, dil=dil, res=r, acc=acc, lens=lens, in_act=in_act, out_act=out_act, alpha=0.5)
        torch.cuda.synchronize()
        outs[name] = y
    assert torch.equal(outs["new"], outs["old"]), (B, T, Cin, N, k, (outs["new"] - outs["old"]).abs().max().item())
print("new == old bit for bit on ragged cases", flush=True)

CASES = (("s0 C256 k3", 256, 256, 3, 1, 8192, False, False), ("s0 C256 k7 res", 256, 256, 7, 1, 8192, True, False), ("s0 C256 k11", 256, 256, 11, 5, 8192, False, False),
         ("s1 C128 k3", 128, 128, 3, 1, 65536, False, False), ("s1 C128 k3 res", 128, 128, 3, 1, 65536, True, False), ("s1 C128 k7", 128, 128, 7, 3, 65536, False, False),
         ("s1 C128 k11 res+acc", 128, 128, 11, 1, 65536, True, True), ("s2 C64 k3", 64, 64, 3, 1, 131072, False, False), ("s2 C64 k7 res", 64, 64, 7, 1, 131072, True, False),
         ("s2 C64 k11 res+acc", 64, 64, 11, 1, 131072, True, True), ("s3 C32 k3", 32, 32, 3, 1, 262144, False, False), ("s3 C32 k7 res", 32, 32, 7, 1, 262144, True, False),
         ("s3 C32 k11 res+acc", 32, 32, 11, 1, 262144, True, True), ("ups1 C256->512 k2", 256, 512, 2, 1, 8192, False, False), ("ups3 C64->32 k2", 64, 32, 2, 1, 131072, False, False),
         ("dec ffn1 k9", 256, 1024, 9, 1, 1024, False, False), ("dec ffn2 k1 res", 1024, 256, 1, 1, 1024, True, False), ("dec qkv", 256, 768, 1, 1, 1024, False, False),
         ("postnet k5", 512, 512, 5, 1, 1024, False, False))
tot = {"new": 0.0, "old": 0.0}
for name, Cin, N, k, dil, T, res, acc in CASES:
    x = torch.randn(16, T, Cin, generator=g).cuda(); r = torch.randn(16, T, N, generator=g).cuda() if res else None
    w = torch.randn(k, Cin, N, generator=g) * (k * Cin) ** -0.5
    wtc = packing.pack_conv_tc(w).cuda(); w = w.cuda()
    y = torch.zeros(16, T, N, device="cuda")
    best = {"new": 1e9, "old": 1e9}
    for rep in range(3):
        for which in ("old", "new"):
            lib = libs[which]
            run(lib, x, w, wtc, y, k=k, dil=dil, res=r, acc=acc, in_act=3); torch.cuda.synchronize()
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); run(lib, x, w, wtc, y, k=k, dil=dil, res=r, acc=acc, in_act=3); e1.record(); torch.cuda.synchronize()
                best[which] = min(best[which], e0.elapsed_time(e1) * 1e3)
    for kx in tot: tot[kx] += best[kx]
    print(f"{name:22s} old {best['old']:7.1f} us | new {best['new']:7.1f} us | {100 * (best['new'] / best['old'] - 1):+6.1f} %", flush=True)
print(f"sum: old {tot['old']:.0f} us, new {tot['new']:.0f} us, {100 * (tot['new'] / tot['old'] - 1):+.1f} %")
