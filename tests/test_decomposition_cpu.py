"""CPU: the op decomposition the CUDA path executes (packing + model.cu's launch sequence, emulated op by op from the
C-ABI contracts) reproduces the oracle.  Tolerances: fp32 re-association only (no reduced precision anywhere)."""
import torch

from fastspeech2_b200 import configs, packing, synth
from oracle import fs2_oracle as O
from tests import emul_cabi as E

CFG = dict(n_head=2, k1=9, k2=1, n_enc=4, n_dec=6, vp_kernel=3, n_postnet=5, post_k=5)


def _pk(sd, multi):
    return packing.pack_acoustic(lambda k: sd[k], 4, 6, 5, multi)


def test_acoustic_free_running_lj(lj_configs):
    pc, mc = lj_configs
    sd = synth.fastspeech2_state_dict(pc, mc, seed=3)
    spk, texts, lens, L = synth.make_batch(3, 28, seed=5, min_len=9)
    ref = O.fastspeech2_forward(sd, spk, texts, lens, L, p_control=1.2, d_control=0.9)
    mel, post, p, e, logd, d, mel_len = E.acoustic_forward(_pk(sd, False), CFG, spk, texts, lens, p_control=1.2, d_control=0.9)
    assert torch.equal(mel_len, ref[9]) and torch.equal(d, ref[5])
    for got, want, tol in ((mel, ref[0], 2e-5), (post, ref[1], 2e-5), (p, ref[2], 5e-6), (e, ref[3], 5e-6), (logd, ref[4], 5e-6)):
        assert (got - want).abs().max() < tol


def test_acoustic_teacher_forced_libri(libri_configs):
    pc, mc = libri_configs
    sd = synth.fastspeech2_state_dict(pc, mc, seed=4)
    spk, texts, lens, L = synth.make_batch(2, 20, seed=6, n_speakers=904, min_len=11)
    g = torch.Generator().manual_seed(1)
    d_t = torch.randint(0, 6, (2, L), generator=g) * (torch.arange(L)[None] < lens[:, None])
    p_t = torch.randn(2, L, generator=g) * 2
    e_t = torch.randn(2, L, generator=g) * 2
    mel_lens = d_t.sum(1)
    T = int(mel_lens.max()) + 3
    ref = O.fastspeech2_forward(sd, spk, texts, lens, L, None, mel_lens, T, p_t, e_t, d_t)
    mel, post, p, e, logd, d, mel_len = E.acoustic_forward(_pk(sd, True), CFG, spk, texts, lens, p_target=p_t, e_target=e_t,
                                                          d_target=d_t.float(), mel_lens=mel_lens, max_mel_len=T)
    assert torch.equal(mel_len, ref[9])
    for got, want in ((mel, ref[0]), (post, ref[1]), (p, ref[2]), (e, ref[3]), (logd, ref[4])):
        assert (got - want).abs().max() < 2e-5


def test_acoustic_frame_level_variances(scratch):
    """config/LJSpeech_paper style: pitch / energy predicted per mel frame after the length regulator (modules.py:139-148)."""
    import copy
    pc, mc = configs.make_configs("LJSpeech", scratch)
    pc = copy.deepcopy(pc)
    pc["preprocessing"]["pitch"]["feature"] = "frame_level"
    pc["preprocessing"]["energy"]["feature"] = "frame_level"
    sd = synth.fastspeech2_state_dict(pc, mc, seed=8)
    spk, texts, lens, L = synth.make_batch(2, 18, seed=9, min_len=11)
    ref = O.fastspeech2_forward(sd, spk, texts, lens, L, p_control=0.9, pitch_level="frame_level", energy_level="frame_level")
    mel, post, p, e, logd, d, mel_len = E.acoustic_forward(_pk(sd, False), CFG, spk, texts, lens, p_control=0.9, pitch_frame=True,
                                                          energy_frame=True)
    assert torch.equal(mel_len, ref[9]) and p.shape == ref[2].shape == mel.shape[:2]
    for got, want in ((mel, ref[0]), (post, ref[1]), (p, ref[2]), (e, ref[3]), (logd, ref[4])):
        assert (got - want).abs().max() < 2e-5


def test_vocoder_decomposition():
    h = configs.HIFIGAN_CONFIG
    sd = synth.hifigan_state_dict(h, seed=2)
    mel = synth.make_mel(2, 12, seed=1)
    want = O.hifigan_forward(sd, mel)
    folded = O.fold_weight_norm(sd)
    pk = packing.pack_vocoder(lambda b: folded[b + ".weight"], lambda b: folded[b + ".bias"], h["upsample_rates"], 12, 3)
    got = E.vocoder_forward(pk, h["upsample_rates"], h["resblock_kernel_sizes"], h["resblock_dilation_sizes"],
                            mel.transpose(1, 2).contiguous())
    assert got.shape == (2, 12 * 256)
    assert (got - want[:, 0]).abs().max() < 2e-5


def test_split_conv_transpose_matches_torch():
    g = torch.Generator().manual_seed(0)
    for u, cin, cout in ((8, 16, 8), (2, 8, 4)):
        w = torch.randn(cin, cout, 2 * u, generator=g)
        x = torch.randn(2, 9, cin, generator=g)
        want = torch.nn.functional.conv_transpose1d(x.transpose(1, 2), w, stride=u, padding=u // 2).transpose(1, 2)
        wa, wb = packing.split_conv_transpose(w, u)
        ya = E.conv1d(x, wa, None, pad_left=1)
        yb = E.conv1d(x, wb, None, pad_left=0)
        got = torch.cat([ya, yb], -1).reshape(2, 9 * u, cout)
        assert (got - want).abs().max() < 1e-5
