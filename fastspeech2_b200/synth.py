"""Deterministic synthetic weights and inputs (there is no network for checkpoints or datasets).

Weights are a pure function of (spec, seed): the same call here, in the golden-vector generator and
on the GPU box yields bit-identical tensors (CPU torch.Generator), which is what lets the committed
fixtures under tests/golden/ carry only inputs and reference OUTPUTS, never the 140 MB of weights.

Magnitudes follow the reference's default initialisers (uniform(+-1/sqrt(fan_in)) for Linear/Conv,
N(0,1) embeddings) with LayerNorm/BatchNorm affine terms and BN running statistics randomised so that
BN folding and the affine paths are exercised.  Raw random init yields ~0.3 frames per phoneme
(SURVEY.md Appendix D), so the duration predictor's last layer is steered (weight x0.05, bias =
log(frames_per_phoneme + 1)) to produce ~1000 mel frames for 128 phonemes, as BASELINE.json's configs ask.
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np
import torch

from .spec import P, fastspeech2_spec, hifigan_spec, read_dataset_files


def sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    """Position table: float64 on the host, cast to fp32 (transformer/Models.py:10-30).
    Vectorised instead of the reference's per-element Python loops; same IEEE operations."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    angle = pos / np.power(10000.0, 2.0 * (j // 2) / d_hid)
    tab = np.empty_like(angle)
    tab[:, 0::2] = np.sin(angle[:, 0::2])
    tab[:, 1::2] = np.cos(angle[:, 1::2])
    return torch.from_numpy(tab).float()


def _fan_in(shape) -> int:
    n = 1
    for s in shape[1:]:
        n *= s
    return max(n, 1)


def _draw(p: P, g: torch.Generator, stats) -> torch.Tensor:
    shape = tuple(p.shape)
    if p.init in ("linear", "conv"):
        b = 1.0 / math.sqrt(_fan_in(shape))
        return (torch.rand(shape, generator=g) * 2 - 1) * b
    if p.init == "bias":
        return (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    if p.init == "ln_w":
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if p.init in ("ln_b", "bn_mean"):
        return 0.1 * torch.randn(shape, generator=g)
    if p.init == "bn_var":
        return 0.5 + torch.rand(shape, generator=g)
    if p.init == "embedding":
        return torch.randn(shape, generator=g)
    if p.init == "embedding_pad0":
        w = torch.randn(shape, generator=g)
        w[0].zero_()
        return w
    if p.init == "sinusoid":
        return sinusoid_table(shape[1], shape[2]).unsqueeze(0)
    if p.init == "pitch_bins":
        return torch.linspace(stats["pitch"][0], stats["pitch"][1], shape[0])
    if p.init == "energy_bins":
        return torch.linspace(stats["energy"][0], stats["energy"][1], shape[0])
    if p.init == "zero":
        return torch.zeros(shape, dtype=torch.long)
    raise ValueError(p.init)


def fastspeech2_state_dict(preprocess_config, model_config, seed: int = 0,
                           frames_per_phoneme: float = 7.8) -> Dict[str, torch.Tensor]:
    stats, _ = read_dataset_files(preprocess_config)
    g = torch.Generator().manual_seed(1000 + seed)
    sd = {p.key: _draw(p, g, stats) for p in fastspeech2_spec(preprocess_config, model_config)}
    ve = model_config["variance_embedding"]
    for name in ("pitch", "energy"):                   # log-spaced edges exactly as the reference builds them (model/modules.py:48-54,:60-67)
        if ve[f"{name}_quantization"] == "log":
            import numpy as np
            sd[f"variance_adaptor.{name}_bins"] = torch.exp(
                torch.linspace(np.log(stats[name][0]), np.log(stats[name][1]), ve["n_bins"] - 1))
    if frames_per_phoneme is not None:
        sd["variance_adaptor.duration_predictor.linear_layer.weight"] *= 0.05
        sd["variance_adaptor.duration_predictor.linear_layer.bias"].fill_(math.log(frames_per_phoneme + 1.0))
    return sd


def hifigan_state_dict(h, seed: int = 0, weight_norm: bool = True, branch_gain: float = 0.6) -> Dict[str, torch.Tensor]:
    """Weight-normed (checkpoint-layout) generator weights with O(1) activations end to end.

    The reference's own init (N(0, 0.01)) gives a ~1e-5 waveform, which would make the 1e-4 parity
    bar meaningless; here each conv is variance-preserving (gain/sqrt(fan_in)) so the waveform is O(0.1)
    like a trained generator's."""
    g = torch.Generator().manual_seed(5000 + seed)
    sd = {}
    for p in hifigan_spec(h, weight_norm=True):
        shape = tuple(p.shape)
        if p.init == "bias":
            sd[p.key] = 0.02 * torch.randn(shape, generator=g)
        elif p.init == "wn_v":
            sd[p.key] = torch.randn(shape, generator=g)
        elif p.init == "wn_g":
            sd[p.key] = torch.ones(shape)          # placeholder, set below once v is known
    for k in [k for k in sd if k.endswith(".weight_v")]:
        v = sd[k]
        base = k[: -len(".weight_v")]
        is_up = base.startswith("ups.")
        # effective fan-in of one output sample
        if is_up:
            fan = v.shape[0] * 2                    # k = 2u  ->  two taps per output phase
            gain = 1.4
        elif base.startswith("resblocks."):
            fan = v.shape[1] * v.shape[2]
            gain = branch_gain * 1.4
        elif base == "conv_post":
            fan = v.shape[1] * v.shape[2]
            gain = 0.2                              # keep tanh out of saturation so errors stay visible
        else:                                       # conv_pre: input mel has std ~2
            fan = v.shape[1] * v.shape[2]
            gain = 0.5
        target_std = gain / math.sqrt(fan)
        # ||v|| over dims != 0; choose g so that the folded weight has elementwise std target_std
        nrm = v.reshape(v.shape[0], -1).norm(dim=1)
        n_el = v[0].numel()
        sd[base + ".weight_g"] = (target_std * math.sqrt(n_el) * (0.8 + 0.4 * torch.rand(nrm.shape, generator=g))
                                  ).reshape(sd[base + ".weight_g"].shape)
    if not weight_norm:
        out = {}
        for k, v in sd.items():
            if k.endswith(".weight_v"):
                gk = sd[k[:-1] + "g"]
                out[k[:-2]] = v * (gk / v.reshape(v.shape[0], -1).norm(dim=1).reshape(gk.shape))
            elif not k.endswith(".weight_g"):
                out[k] = v
        return out
    return sd


def make_batch(batch: int, max_len: int, seed: int = 0, n_speakers: int = 1, min_len: int | None = None):
    """Synthetic phoneme batch in the layout synthesize.py:203-210 feeds the model:
    (speakers i64[B], texts i64[B,L] with 0 = PAD, src_lens i64[B], max_src_len int)."""
    g = torch.Generator().manual_seed(9000 + seed)
    if min_len is None or min_len >= max_len:
        lens = torch.full((batch,), max_len, dtype=torch.long)
    else:
        lens = torch.randint(min_len, max_len + 1, (batch,), generator=g)
        lens[0] = max_len
    L = int(lens.max())
    texts = torch.randint(1, 361, (batch, L), generator=g)
    texts = texts * (torch.arange(L)[None, :] < lens[:, None])
    speakers = torch.randint(0, n_speakers, (batch,), generator=g) if n_speakers > 1 else torch.zeros(batch, dtype=torch.long)
    return speakers, texts, lens, L


def make_mel(batch: int, frames: int, seed: int = 0) -> torch.Tensor:
    """Synthetic log-mel input for vocoder-only tests: N(-5, 2) clipped to [-11.5, 2]."""
    g = torch.Generator().manual_seed(7000 + seed)
    return (torch.randn(batch, 80, frames, generator=g) * 2.0 - 5.0).clamp_(-11.5, 2.0)
