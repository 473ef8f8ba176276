"""GPU: the two drop-in modules against the CPU oracle on identical inputs and weights.

Protocol (hard part 2): (A) free-running -- continuous predictions are compared and the discrete
decisions (durations, hence mel_lens) must match exactly; (B) teacher-forced with the oracle's own decisions so a
boundary flip cannot hide or fake a mel error.  Bars from BASELINE.json: mel 1e-3, waveform 1e-4 max-abs."""
import pytest
import torch

from fastspeech2_b200 import configs, synth
from fastspeech2_b200.hifigan import AttrDict, Generator
from fastspeech2_b200.model import FastSpeech2
from oracle import fs2_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
MEL_TOL, WAV_TOL = 1e-3, 1e-4


def _model(cfgs, seed):
    pc, mc = cfgs
    sd = synth.fastspeech2_state_dict(pc, mc, seed=seed)
    m = FastSpeech2(pc, mc)
    m.load_state_dict(sd)
    return m.to(DEV).eval(), sd


def _cmp(out, ref):
    errs = {}
    for i, name in ((0, "mel"), (1, "postnet"), (2, "pitch"), (3, "energy"), (4, "logd")):
        errs[name] = (out[i].cpu() - ref[i]).abs().max().item()
    return errs


@pytest.mark.parametrize("B,L,min_len", [(1, 24, None), (3, 40, 17), (16, 128, None)])
def test_fastspeech2_free_running_lj(lj_configs, B, L, min_len, parity_log):
    m, sd = _model(lj_configs, seed=1)
    spk, texts, lens, Lm = synth.make_batch(B, L, seed=2, min_len=min_len)
    ref = O.fastspeech2_forward(sd, spk, texts, lens, Lm)
    out = m(spk.to(DEV), texts.to(DEV), lens.to(DEV), Lm)
    assert torch.equal(out[5].cpu(), ref[5]), "duration decisions differ"
    assert torch.equal(out[9].cpu(), ref[9]) and torch.equal(out[6].cpu(), ref[6]) and torch.equal(out[7].cpu(), ref[7])
    e = _cmp(out, ref)
    parity_log(f"fs2_free_running_lj_B{B}_L{L}", **e)
    assert e["mel"] < MEL_TOL and e["postnet"] < MEL_TOL and max(e["pitch"], e["energy"], e["logd"]) < 1e-4, e


def test_fastspeech2_multispeaker_controls(libri_configs, parity_log):
    m, sd = _model(libri_configs, seed=3)
    spk, texts, lens, Lm = synth.make_batch(5, 64, seed=4, n_speakers=904, min_len=20)
    kw = dict(p_control=1.15, e_control=0.8, d_control=1.3)
    ref = O.fastspeech2_forward(sd, spk, texts, lens, Lm, **kw)
    out = m(spk.to(DEV), texts.to(DEV), lens.to(DEV), Lm, **kw)
    assert torch.equal(out[5].cpu(), ref[5]) and torch.equal(out[9].cpu(), ref[9])
    e = _cmp(out, ref)
    parity_log("fs2_multispeaker_controls_B5_L64", **e)
    assert e["mel"] < MEL_TOL and e["postnet"] < MEL_TOL, e


def test_fastspeech2_ragged_multispeaker_long(libri_configs, parity_log):
    """BASELINE.json configs[3] in miniature: LibriTTS multi-speaker, mixed 64-256 phonemes with padding masks -> T up to ~2000
    (position table beyond max_seq_len, decoder attention with 2048 padded keys)."""
    m, sd = _model(libri_configs, seed=13)
    spk, texts, lens, Lm = synth.make_batch(6, 256, seed=14, n_speakers=904, min_len=64)
    ref = O.fastspeech2_forward(sd, spk, texts, lens, Lm)
    out = m(spk.to(DEV), texts.to(DEV), lens.to(DEV), Lm)
    assert torch.equal(out[5].cpu(), ref[5]) and torch.equal(out[9].cpu(), ref[9]) and torch.equal(out[7].cpu(), ref[7])
    assert int(ref[9].max()) > 1500
    e = _cmp(out, ref)
    parity_log("fs2_ragged_libri_B6_L64-256", **e, tmax=int(ref[9].max()))
    assert e["mel"] < MEL_TOL and e["postnet"] < MEL_TOL, e


def test_fastspeech2_full_size_tensor_core_vs_exact_path(libri_configs, parity_log):
    """BASELINE.json configs[3] at FULL size (LibriTTS, B = 64, 64-256 phonemes, Tmax ~ 2000), where the CPU oracle would take
    minutes: the tcgen05 decoder / PostNet must agree with the independently written exact-fp32 kernels (which the tests above
    pin to the oracle) on identical decisions, and padded rows must follow the reference's padding semantics."""
    pc, mc = libri_configs
    sd = synth.fastspeech2_state_dict(pc, mc, seed=21)
    fast, exact = FastSpeech2(pc, mc), FastSpeech2(pc, mc)
    fast.load_state_dict(sd); exact.load_state_dict(sd)
    exact.tc_mask = 0
    fast, exact = fast.to(DEV).eval(), exact.to(DEV).eval()
    spk, texts, lens, Lm = synth.make_batch(64, 256, seed=22, n_speakers=904, min_len=64)
    a = fast(spk.to(DEV), texts.to(DEV), lens.to(DEV), Lm)
    b = exact(spk.to(DEV), texts.to(DEV), lens.to(DEV), Lm)
    torch.cuda.synchronize()
    assert torch.equal(a[5], b[5]) and torch.equal(a[9], b[9]) and torch.equal(a[7], b[7])
    assert int(a[9].max()) > 1800 and a[0].shape[0] == 64
    e = {"mel": (a[0] - b[0]).abs().max().item(), "postnet": (a[1] - b[1]).abs().max().item(), "tmax": int(a[9].max()),
         "frames": int(a[9].sum())}
    parity_log("fs2_full_size_libri_B64_tc_vs_exact", **e)
    assert e["mel"] < 2e-4 and e["postnet"] < 2e-4, e          # each path is within ~3e-5 of the oracle at this length; bar 1e-3
    # padded mel rows equal mel_linear.bias exactly (decoder output is zeroed there)
    bias = sd["mel_linear.bias"].to(DEV)
    pad = a[7]                                             # True = padded frame
    assert pad.any() and (a[0][pad] - bias).abs().max().item() < 1e-6
    assert torch.isfinite(a[1]).all()


def test_fastspeech2_frame_level_variances(scratch, parity_log):
    """pitch / energy feature = frame_level (config/LJSpeech_paper): predictors run on the expanded sequence."""
    import copy
    pc, mc = configs.make_configs("LJSpeech", scratch)
    pc = copy.deepcopy(pc)
    pc["preprocessing"]["pitch"]["feature"] = "frame_level"
    pc["preprocessing"]["energy"]["feature"] = "frame_level"
    sd = synth.fastspeech2_state_dict(pc, mc, seed=15)
    m = FastSpeech2(pc, mc); m.load_state_dict(sd); m = m.to(DEV).eval()
    spk, texts, lens, Lm = synth.make_batch(3, 40, seed=16, min_len=21)
    ref = O.fastspeech2_forward(sd, spk, texts, lens, Lm, p_control=1.1, pitch_level="frame_level", energy_level="frame_level")
    out = m(spk.to(DEV), texts.to(DEV), lens.to(DEV), Lm, p_control=1.1)
    assert torch.equal(out[9].cpu(), ref[9]) and out[2].shape == ref[2].shape
    e = _cmp(out, ref)
    parity_log("fs2_frame_level_B3_L40", **e)
    assert e["mel"] < MEL_TOL and e["postnet"] < MEL_TOL and max(e["pitch"], e["energy"]) < 1e-4, e


def test_fastspeech2_teacher_forced(lj_configs, parity_log):
    m, sd = _model(lj_configs, seed=5)
    spk, texts, lens, Lm = synth.make_batch(4, 48, seed=6, min_len=15)
    free = O.fastspeech2_forward(sd, spk, texts, lens, Lm)
    d_t, mel_lens = free[5].long(), free[9]
    T = int(mel_lens.max())
    ref = O.fastspeech2_forward(sd, spk, texts, lens, Lm, None, mel_lens, T, free[2], free[3], d_t)
    out = m(spk.to(DEV), texts.to(DEV), lens.to(DEV), Lm, None, mel_lens.to(DEV), T, free[2].to(DEV), free[3].to(DEV), d_t.to(DEV))
    e = _cmp(out, ref)
    parity_log("fs2_teacher_forced_B4_L48", **e)
    assert e["mel"] < MEL_TOL and e["postnet"] < MEL_TOL, e
    assert torch.equal(out[9].cpu(), ref[9])


def test_fastspeech2_long_sequence_position_table(lj_configs):
    """T > max_seq_len (1000): eval mode recomputes the sinusoid table and never truncates (Models.py:145-152)."""
    pc, mc = lj_configs
    sd = synth.fastspeech2_state_dict(pc, mc, seed=7, frames_per_phoneme=11.0)
    m = FastSpeech2(pc, mc); m.load_state_dict(sd); m = m.to(DEV).eval()
    spk, texts, lens, Lm = synth.make_batch(2, 120, seed=8, min_len=60)
    ref = O.fastspeech2_forward(sd, spk, texts, lens, Lm)
    assert int(ref[9].max()) > 1000
    out = m(spk.to(DEV), texts.to(DEV), lens.to(DEV), Lm)
    assert torch.equal(out[9].cpu(), ref[9])
    e = _cmp(out, ref)
    assert e["mel"] < MEL_TOL and e["postnet"] < MEL_TOL, e


def test_fastspeech2_golden_vs_reference(lj_configs, libri_configs):
    """Committed outputs of the UNMODIFIED reference (tests/golden/, made by oracle/gen_golden.py)."""
    import numpy as np, os
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    for name, cfgs in (("fs2_lj", lj_configs), ("fs2_libri", libri_configs)):
        z = np.load(os.path.join(gdir, name + ".npz"))
        m, sd = _model(cfgs, seed=int(z["seed"]))
        t = lambda k: torch.from_numpy(z[k])
        out = m(t("speakers").to(DEV), t("texts").to(DEV), t("src_lens").to(DEV), int(z["max_src_len"]),
                p_control=float(z["p_control"]), e_control=float(z["e_control"]), d_control=float(z["d_control"]))
        assert torch.equal(out[9].cpu(), t("mel_lens")) and torch.equal(out[5].cpu(), t("d_rounded"))
        assert (out[0].cpu() - t("mel")).abs().max() < MEL_TOL
        assert (out[1].cpu() - t("postnet_mel")).abs().max() < MEL_TOL


def _generator(seed):
    h = AttrDict(configs.HIFIGAN_CONFIG)
    sd = synth.hifigan_state_dict(h, seed=seed)
    gen = Generator(h)
    gen.load_state_dict(sd)
    gen.eval()
    gen.remove_weight_norm()
    return gen.to(DEV), sd


@pytest.mark.parametrize("B,T", [(1, 7), (2, 50), (3, 129)])
def test_hifigan_vs_oracle(B, T, parity_log):
    gen, sd = _generator(seed=1)
    mel = synth.make_mel(B, T, seed=2)
    want = O.hifigan_forward(sd, mel)
    got = gen(mel.to(DEV))
    assert got.shape == want.shape
    want64 = O.hifigan_forward(sd, mel, dtype=torch.float64)
    gen.use_tensor_cores = False; gen._invalidate()
    simt = gen(mel.to(DEV)).cpu()
    gen.use_tensor_cores = True; gen._invalidate()
    parity_log(f"hifigan_B{B}_T{T}", wav_vs_oracle32=(got.cpu() - want).abs().max(), wav_vs_oracle64=(got.cpu().double() - want64).abs().max(),
               simt_vs_oracle32=(simt - want).abs().max(), oracle32_vs_64=(want.double() - want64).abs().max(), peak=want.abs().max())
    assert (got.cpu() - want).abs().max() < WAV_TOL
    # the usual caller passes a transposed channels-last view (utils/tools.py:202)
    view = mel.transpose(1, 2).contiguous().to(DEV).transpose(1, 2)
    assert (gen(view).cpu() - want).abs().max() < WAV_TOL


def test_hifigan_full_size_tensor_core_vs_exact_path(parity_log):
    """BASELINE.json configs[2]'s vocoder at FULL size (B = 16 x 1012 frames -> 16 x 259072 samples; the CPU oracle needs about a
    minute per utterance): the tcgen05 path against the exact-fp32 kernels that the small cases above pin to the oracle."""
    gen, _ = _generator(seed=5)
    mel = synth.make_mel(16, 1012, seed=6).to(DEV)
    got = gen(mel)
    gen.use_tensor_cores = False; gen._invalidate()
    exact = gen(mel)
    gen.use_tensor_cores = True; gen._invalidate()
    torch.cuda.synchronize()
    assert got.shape == (16, 1, 1012 * 256) and torch.isfinite(got).all()
    err, peak = (got - exact).abs().max().item(), exact.abs().max().item()
    parity_log("hifigan_full_size_B16_T1012_tc_vs_exact", wav_tc_vs_exact=err, peak=peak)
    assert err < WAV_TOL and peak > 0.3, (err, peak)


def test_hifigan_golden_vs_reference():
    import numpy as np, os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "hifigan.npz"))
    gen, _ = _generator(seed=int(z["seed"]))
    got = gen(torch.from_numpy(z["mel"]).to(DEV))
    assert (got.cpu() - torch.from_numpy(z["wav"])).abs().max() < WAV_TOL


def test_end_to_end_wav(lj_configs):
    m, sd = _model(lj_configs, seed=9)
    gen, hsd = _generator(seed=3)
    spk, texts, lens, Lm = synth.make_batch(2, 32, seed=10, min_len=20)
    ref = O.fastspeech2_forward(sd, spk, texts, lens, Lm)
    out = m(spk.to(DEV), texts.to(DEV), lens.to(DEV), Lm)
    assert torch.equal(out[9].cpu(), ref[9])
    want = O.hifigan_forward(hsd, ref[1].transpose(1, 2))
    got = gen(out[1].transpose(1, 2))
    # waveform bar applies to the vocoder given the same mel; end to end the mel error (<=1e-3) propagates, so compare both ways
    same_mel = gen(ref[1].to(DEV).transpose(1, 2))
    assert (same_mel.cpu() - want).abs().max() < WAV_TOL
    assert (got.cpu() - want).abs().max() < 5e-3
