"""Parameter layout (state_dict key contract) of the two modules on the hot path.

The drop-in boundary is the reference's ``nn.Module`` pair; a replacement must expose the same
``state_dict`` keys and shapes so reference checkpoints load unchanged (utils/model.py:15-21, :62-66).
Key names/shapes follow the reference constructors:
  model/fastspeech2.py:16-41, model/modules.py:20-78 and :200-240, transformer/Models.py:36-71 and :106-137,
  transformer/SubLayers.py:11-27 and :63-83, transformer/Layers.py:67-127, hifigan/models.py:20-94 and :112-147.
Only the names are shared with the reference; construction code here is table-driven.
"""
from __future__ import annotations

import json
import os
from typing import List, NamedTuple, Tuple

N_SRC_VOCAB = 361          # len(text.symbols.symbols) + 1  (transformer/Models.py:40; 360 symbols)


class P(NamedTuple):
    key: str
    shape: Tuple[int, ...]
    kind: str               # "param" | "frozen" (Parameter, requires_grad=False) | "buffer" | "buffer_long"
    init: str               # initialiser tag understood by synth.py


def _fft_block(prefix: str, d: int, n_head: int, d_inner: int, ks) -> List[P]:
    dk = d // n_head
    out = []
    for nm in ("w_qs", "w_ks", "w_vs"):
        out += [P(f"{prefix}.slf_attn.{nm}.weight", (n_head * dk, d), "param", "linear"),
                P(f"{prefix}.slf_attn.{nm}.bias", (n_head * dk,), "param", "bias")]
    out += [P(f"{prefix}.slf_attn.layer_norm.weight", (d,), "param", "ln_w"),
            P(f"{prefix}.slf_attn.layer_norm.bias", (d,), "param", "ln_b"),
            P(f"{prefix}.slf_attn.fc.weight", (d, n_head * dk), "param", "linear"),
            P(f"{prefix}.slf_attn.fc.bias", (d,), "param", "bias"),
            P(f"{prefix}.pos_ffn.w_1.weight", (d_inner, d, ks[0]), "param", "conv"),
            P(f"{prefix}.pos_ffn.w_1.bias", (d_inner,), "param", "bias"),
            P(f"{prefix}.pos_ffn.w_2.weight", (d, d_inner, ks[1]), "param", "conv"),
            P(f"{prefix}.pos_ffn.w_2.bias", (d,), "param", "bias"),
            P(f"{prefix}.pos_ffn.layer_norm.weight", (d,), "param", "ln_w"),
            P(f"{prefix}.pos_ffn.layer_norm.bias", (d,), "param", "ln_b")]
    return out


def _predictor(prefix: str, d_in: int, d_f: int, k: int) -> List[P]:
    return [P(f"{prefix}.conv_layer.conv1d_1.conv.weight", (d_f, d_in, k), "param", "conv"),
            P(f"{prefix}.conv_layer.conv1d_1.conv.bias", (d_f,), "param", "bias"),
            P(f"{prefix}.conv_layer.layer_norm_1.weight", (d_f,), "param", "ln_w"),
            P(f"{prefix}.conv_layer.layer_norm_1.bias", (d_f,), "param", "ln_b"),
            P(f"{prefix}.conv_layer.conv1d_2.conv.weight", (d_f, d_f, k), "param", "conv"),
            P(f"{prefix}.conv_layer.conv1d_2.conv.bias", (d_f,), "param", "bias"),
            P(f"{prefix}.conv_layer.layer_norm_2.weight", (d_f,), "param", "ln_w"),
            P(f"{prefix}.conv_layer.layer_norm_2.bias", (d_f,), "param", "ln_b"),
            P(f"{prefix}.linear_layer.weight", (1, d_f), "param", "linear"),
            P(f"{prefix}.linear_layer.bias", (1,), "param", "bias")]


def read_dataset_files(preprocess_config):
    root = preprocess_config["path"]["preprocessed_path"]
    with open(os.path.join(root, "stats.json")) as f:
        stats = json.load(f)
    n_speaker = 0
    sp = os.path.join(root, "speakers.json")
    if os.path.exists(sp):
        with open(sp) as f:
            n_speaker = len(json.load(f))
    return stats, n_speaker


def fastspeech2_spec(preprocess_config, model_config) -> List[P]:
    tr = model_config["transformer"]
    d_enc, d_dec = tr["encoder_hidden"], tr["decoder_hidden"]
    n_pos = model_config["max_seq_len"] + 1
    n_mel = preprocess_config["preprocessing"]["mel"]["n_mel_channels"]
    vp = model_config["variance_predictor"]
    n_bins = model_config["variance_embedding"]["n_bins"]
    spec = [P("encoder.position_enc", (1, n_pos, d_enc), "frozen", "sinusoid"),
            P("encoder.src_word_emb.weight", (N_SRC_VOCAB, d_enc), "param", "embedding_pad0")]
    for i in range(tr["encoder_layer"]):
        spec += _fft_block(f"encoder.layer_stack.{i}", d_enc, tr["encoder_head"], tr["conv_filter_size"],
                           tr["conv_kernel_size"])
    spec += [P("variance_adaptor.pitch_bins", (n_bins - 1,), "frozen", "pitch_bins"),
             P("variance_adaptor.energy_bins", (n_bins - 1,), "frozen", "energy_bins")]
    for nm in ("duration", "pitch", "energy"):
        spec += _predictor(f"variance_adaptor.{nm}_predictor", d_enc, vp["filter_size"], vp["kernel_size"])
    spec += [P("variance_adaptor.pitch_embedding.weight", (n_bins, d_enc), "param", "embedding"),
             P("variance_adaptor.energy_embedding.weight", (n_bins, d_enc), "param", "embedding"),
             P("decoder.position_enc", (1, n_pos, d_dec), "frozen", "sinusoid")]
    for i in range(tr["decoder_layer"]):
        spec += _fft_block(f"decoder.layer_stack.{i}", d_dec, tr["decoder_head"], tr["conv_filter_size"],
                           tr["conv_kernel_size"])
    spec += [P("mel_linear.weight", (n_mel, d_dec), "param", "linear"),
             P("mel_linear.bias", (n_mel,), "param", "bias")]
    chans = [n_mel, 512, 512, 512, 512, n_mel]              # PostNet() takes no args: transformer/Layers.py:72-78
    for i in range(5):
        p = f"postnet.convolutions.{i}"
        spec += [P(f"{p}.0.conv.weight", (chans[i + 1], chans[i], 5), "param", "conv"),
                 P(f"{p}.0.conv.bias", (chans[i + 1],), "param", "bias"),
                 P(f"{p}.1.weight", (chans[i + 1],), "param", "ln_w"),
                 P(f"{p}.1.bias", (chans[i + 1],), "param", "ln_b"),
                 P(f"{p}.1.running_mean", (chans[i + 1],), "buffer", "bn_mean"),
                 P(f"{p}.1.running_var", (chans[i + 1],), "buffer", "bn_var"),
                 P(f"{p}.1.num_batches_tracked", (), "buffer_long", "zero")]
    if model_config["multi_speaker"]:
        _, n_speaker = read_dataset_files(preprocess_config)
        spec.append(P("speaker_emb.weight", (n_speaker, d_enc), "param", "embedding"))
    return spec


def hifigan_spec(h, weight_norm: bool = True) -> List[P]:
    """Keys of hifigan.Generator (resblock '1').  With ``weight_norm`` the conv weights appear as
    ``weight_g`` / ``weight_v`` (the checkpoint layout)."""
    def conv(prefix, shape):
        if weight_norm:
            g = (shape[0],) + (1,) * (len(shape) - 1)
            return [P(prefix + ".bias", (shape[1] if "ups." in prefix else shape[0],), "param", "bias"),
                    P(prefix + ".weight_g", g, "param", "wn_g"),
                    P(prefix + ".weight_v", shape, "param", "wn_v")]
        return [P(prefix + ".weight", shape, "param", "conv"),
                P(prefix + ".bias", (shape[1] if "ups." in prefix else shape[0],), "param", "bias")]

    c0 = h["upsample_initial_channel"]
    spec = conv("conv_pre", (c0, 80, 7))
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        spec += conv(f"ups.{i}", (c0 // 2 ** i, c0 // 2 ** (i + 1), k))
    nk = len(h["resblock_kernel_sizes"])
    ch = c0
    for i in range(len(h["upsample_rates"])):
        ch = c0 // 2 ** (i + 1)
        for j, k in enumerate(h["resblock_kernel_sizes"]):
            for grp in ("convs1", "convs2"):
                for m in range(len(h["resblock_dilation_sizes"][j])):
                    spec += conv(f"resblocks.{i * nk + j}.{grp}.{m}", (ch, ch, k))
    spec += conv("conv_post", (1, ch, 7))
    return spec
