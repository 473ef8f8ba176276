"""CPU: the C-ABI library builds/loads without a GPU and exports every symbol include/fs2b200.h declares."""
import ctypes
import os
import re

import pytest

from fastspeech2_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "fs2b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fs2_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_loads():
    path = build.build()
    assert os.path.exists(path)
    h = _lib.lib()
    assert h.fs2_abi_version() == _lib.ABI_VERSION
    assert b"sm_100a" in h.fs2_build_info()


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert len(names) >= 19
    h = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(h, n), f"{n} declared in fs2b200.h but not exported"
        assert n in _lib.EXPORTS, f"{n} has no ctypes binding"


def test_struct_sizes_match_the_header():
    h = _lib.lib()
    table = {0: _lib.Conv1dArgs, 1: _lib.LayerNormArgs, 2: _lib.AttentionArgs, 3: _lib.EmbedArgs, 4: _lib.RowBiasArgs,
             5: _lib.VarianceHeadArgs, 6: _lib.DurationsArgs, 7: _lib.LengthRegulateArgs, 8: _lib.ConvPostArgs,
             9: _lib.AcousticModel, 10: _lib.EncodeArgs, 11: _lib.DecodeArgs, 12: _lib.VocoderModel, 13: _lib.VocoderArgs}
    for i, cls in table.items():
        assert h.fs2_struct_size(i) == ctypes.sizeof(cls), cls.__name__


def test_host_side_argument_checks_need_no_gpu():
    h = _lib.lib()
    assert h.fs2_conv1d(None, None) == -1                       # FS2_ERR_ARG
    a = _lib.Conv1dArgs(x=16, w=16, y=16, B=1, T=4, Cin=24, N=16, taps=1)
    assert h.fs2_conv1d(ctypes.byref(a), None) == -2             # Cin % 16 -> FS2_ERR_UNSUPPORTED
    m = _lib.AcousticModel(d_model=256, n_head=2, d_inner=1024, k1=9, k2=1, n_enc=4, n_dec=6, n_mel=80, vp_filter=256,
                           vp_kernel=3, n_postnet=5)
    for i, c in enumerate([512, 512, 512, 512, 80]):
        m.post_cout[i] = c
    need = h.fs2_decode_workspace_bytes(ctypes.byref(m), 16, 1024)
    assert need > 16 * 1024 * (256 * 4 + 768 + 1024) * 4
    assert h.fs2_encode_workspace_bytes(ctypes.byref(m), 0, 5) == 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fastspeech2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_cpu_tensors_fail_loudly(lj_configs):
    import torch
    from fastspeech2_b200 import synth
    from fastspeech2_b200.model import FastSpeech2
    pc, mc = lj_configs
    m = FastSpeech2(pc, mc).eval()
    spk, texts, lens, L = synth.make_batch(1, 8)
    with pytest.raises(_lib.Fs2Error):
        m(spk, texts, lens, L)
