"""A/B (one process): programmatic dependent launch of the conv layers: 0 = off, 1 = short launches only, 2 = all (debug switch fs2_debug_set_tc_pdl).
Times the bench workload's two forwards and checks that the outputs are bit-identical."""
import ctypes, os, sys, tempfile, contextlib, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastspeech2_b200 import _lib, configs, synth
from fastspeech2_b200.hifigan import AttrDict, Generator
from fastspeech2_b200.model import FastSpeech2
lib = _lib.lib()
lib.fs2_debug_set_tc_pdl.argtypes = [ctypes.c_int]
dev = torch.device("cuda:0")
pc, mc = configs.make_configs("LJSpeech", tempfile.mkdtemp())
model = FastSpeech2(pc, mc); model.load_state_dict(synth.fastspeech2_state_dict(pc, mc, seed=0)); model = model.to(dev).eval()
voc = Generator(AttrDict(configs.HIFIGAN_CONFIG)); voc.load_state_dict(synth.hifigan_state_dict(configs.HIFIGAN_CONFIG, seed=0)); voc.eval()
with contextlib.redirect_stdout(io.StringIO()):
    voc.remove_weight_norm()
voc.to(dev)
spk, texts, lens, L = synth.make_batch(16, 128, seed=0)
spk, texts, lens = spk.to(dev), texts.to(dev), lens.to(dev)


def step():
    out = model(spk, texts, lens, L)
    return out[1], voc(out[1].transpose(1, 2))


def one(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), r

import statistics
for _ in range(3): step()
ref = None
res = {m: {"step": [], "fs2": [], "voc": []} for m in (0, 1, 2)}
mel_in = step()[0].transpose(1, 2)
for i in range(45):                      # modes interleaved iteration by iteration: clock / power drift hits all of them alike
    m = i % 3
    lib.fs2_debug_set_tc_pdl(m)
    ms, (mel, wav) = one(step)
    if ref is None: ref = (mel.clone(), wav.clone())
    assert torch.equal(ref[0], mel) and torch.equal(ref[1], wav)
    res[m]["step"].append(ms)
    res[m]["fs2"].append(one(lambda: model(spk, texts, lens, L))[0])
    res[m]["voc"].append(one(lambda: voc(mel_in))[0])
for m in (0, 1, 2):
    print(f"pdl={m}: " + " | ".join(f"{k} median {statistics.median(v):7.3f} min {min(v):7.3f} ms" for k, v in res[m].items()), flush=True)
print("outputs bit-identical across modes")
