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


def test_conv_post_sliding_window_schedule():
    """The work decomposition of conv_post_c32_kernel (rowwise.cu), emulated lane for lane: 8 lanes per time row (4 channels each), groups
    of 18 x 7 input rows = 120 output rows + a 6-row halo, 7 sliding accumulators per lane (accumulator k of row r belongs to output
    r - 3 + k and takes tap 6 - k), an 8-lane sum when an output has seen its last row, stores in runs of 8 samples.  Must equal
    lrelu -> Conv1d(32, 1, 7, padding=3) -> tanh (hifigan/models.py:161-163) for lengths around the group and store-run boundaries,
    with every sample written exactly once."""
    import numpy as np
    TAPS, PAD, BLOCKS = 7, 3, 18
    ROWS = BLOCKS * TAPS - 2 * PAD
    rng = np.random.default_rng(0)
    for T in (1, 5, 7, 8, 9, 119, 120, 121, 127, 128, 250, 963):
        x = rng.standard_normal((T, 32))
        w = rng.standard_normal((TAPS, 32)) * 0.1
        xa = np.where(x > 0, x, 0.01 * x)
        want = np.array([np.tanh(0.05 + sum((xa[t + j - PAD] * w[j]).sum() for j in range(TAPS) if 0 <= t + j - PAD < T)) for t in range(T)])
        out = np.full(T, np.nan)
        for g in range((T + ROWS - 1) // ROWS):
            t0 = g * ROWS
            tend = min(t0 + ROWS, T)
            s, keep = np.zeros((8, TAPS)), np.zeros(8)
            for blk in range(BLOCKS):
                for i in range(TAPS):
                    r = t0 - PAD + blk * TAPS + i
                    v = xa[r].reshape(8, 4) if 0 <= r < T else np.zeros((8, 4))
                    for k in range(TAPS):
                        s[:, k] += (v * w[TAPS - 1 - k].reshape(8, 4)).sum(1)
                    tot = s[:, 0].sum()
                    s[:, :-1] = s[:, 1:].copy()
                    s[:, -1] = 0.0
                    t = r - PAD
                    if t0 <= t < tend:
                        o = (t - t0) & 7
                        keep[o] = tot
                        if o == 7 or t == tend - 1:
                            for sub in range(o + 1):
                                assert np.isnan(out[t - o + sub])
                                out[t - o + sub] = np.tanh(keep[sub] + 0.05)
        assert not np.isnan(out).any() and np.abs(out - want).max() < 1e-12, T
