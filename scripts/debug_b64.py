import sys, os, tempfile, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastspeech2_b200 import configs, synth, _lib as L
from fastspeech2_b200.model import FastSpeech2
pc, mc = configs.make_configs("LJSpeech", tempfile.mkdtemp())
sd = synth.fastspeech2_state_dict(pc, mc, seed=0)
DEV = "cuda"
def run(tc_mask, B):
    m = FastSpeech2(pc, mc); m.load_state_dict(sd); m.tc_mask = tc_mask; m = m.to(DEV).eval()
    spk, texts, lens, Lm = synth.make_batch(B, 128, seed=3)
    out = m(spk.to(DEV), texts.to(DEV), lens.to(DEV), Lm)
    torch.cuda.synchronize()
    return out
for B in (16, 32, 48, 64):
    exact = run(0, B)
    for name, mask in (("tc_split3", L.TC_DECODER | L.TC_POSTNET), ("tc_f8", L.TC_DECODER | L.TC_POSTNET | L.TC_DECODER_F8 | L.TC_POSTNET_F8),
                       ("postnet_only", L.TC_POSTNET), ("decoder_only", L.TC_DECODER)):
        o = run(mask, B)
        per_b = (o[0] - exact[0]).abs().amax(dim=(1, 2))
        bad = (per_b > 1e-3).nonzero().flatten().tolist()
        print(f"B={B} {name}: T={o[0].shape[1]} max err {per_b.max().item():.3e} bad utterances {bad[:20]} ({len(bad)})", flush=True)
        if bad:
            b = bad[0]
            row_err = (o[0][b] - exact[0][b]).abs().amax(dim=1)
            nz = (row_err > 1e-3).nonzero().flatten()
            print("   first bad utt", b, "bad rows", nz[:8].tolist(), "...", nz[-8:].tolist(), "count", nz.numel(), "mel_len", int(o[9][b]))
