"""One bench step under cudaProfilerStart/Stop, for ncu (--profile-from-start off).  Usage: [--vocoder-only] [--batch B]"""
import os, sys, tempfile, contextlib, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastspeech2_b200 import configs, synth
from fastspeech2_b200.hifigan import AttrDict, Generator
from fastspeech2_b200.model import FastSpeech2

B = int(sys.argv[sys.argv.index("--batch") + 1]) if "--batch" in sys.argv else 16
dev = torch.device("cuda:0")
pc, mc = configs.make_configs("LJSpeech", tempfile.mkdtemp())
model = FastSpeech2(pc, mc); model.load_state_dict(synth.fastspeech2_state_dict(pc, mc, seed=0)); model = model.to(dev).eval()
voc = Generator(AttrDict(configs.HIFIGAN_CONFIG)); voc.load_state_dict(synth.hifigan_state_dict(configs.HIFIGAN_CONFIG, seed=0)); voc.eval()
with contextlib.redirect_stdout(io.StringIO()):
    voc.remove_weight_norm()
voc.to(dev)
spk, texts, lens, L = synth.make_batch(B, 128, seed=0)
spk, texts, lens = spk.to(dev), texts.to(dev), lens.to(dev)

def step():
    if "--vocoder-only" in sys.argv:
        return voc(step.mel)
    out = model(spk, texts, lens, L)
    return voc(out[1].transpose(1, 2))

step.mel = model(spk, texts, lens, L)[1].transpose(1, 2)
step(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
