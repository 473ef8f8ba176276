"""The shipped HiFi-GAN generator checkpoints as a travelling fixture.  TEST INFRASTRUCTURE (oracle/ rules apply).

The reference ships its real vocoder weights zipped (hifigan/generator_{LJSpeech,universal}.pth.tar.zip, loaded by
utils/model.py:58-69 as ckpt["generator"]).  /root/reference does not exist on the GPU box, so `make()` -- run by
__graft_entry__.build() wherever the reference tree is present -- unzips one into oracle/_ref/ (git-ignored, NOT gpurun-ignored:
it travels with the snapshot like the built .so).  The matching reference OUTPUTS are committed: tests/golden/hifigan_real_*.npz
(oracle/gen_golden.py).  Tests skip when the fixture is absent.
"""
from __future__ import annotations

import io
import os
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("FS2_REFERENCE", "/root/reference")


def fixture_path(name: str = "LJSpeech") -> str:
    return os.path.join(REF_DIR, f"hifigan_generator_{name}.pt")


def source_zip(name: str = "LJSpeech") -> str:
    return os.path.join(REFERENCE_ROOT, "hifigan", f"generator_{name}.pth.tar.zip")


def read_reference_checkpoint(name: str = "LJSpeech"):
    """state_dict ckpt["generator"] straight from the reference's zip (weight-normed keys, 234 tensors)."""
    import torch
    inner = f"generator_{name}.pth.tar"
    with zipfile.ZipFile(source_zip(name)) as z:
        return torch.load(io.BytesIO(z.read(inner)), map_location="cpu")["generator"]


def make(name: str = "LJSpeech", force: bool = False) -> str | None:
    """Write oracle/_ref/hifigan_generator_<name>.pt from the reference tree; None when the tree is absent."""
    import torch
    dst = fixture_path(name)
    if os.path.exists(dst) and not force:
        return dst
    if not os.path.exists(source_zip(name)):
        return None
    os.makedirs(REF_DIR, exist_ok=True)
    sd = {k: v.contiguous() for k, v in read_reference_checkpoint(name).items()}
    torch.save({"generator": sd}, dst)
    return dst


def load(name: str = "LJSpeech"):
    """The fixture's state_dict, or None when it has not been made (GPU box without a snapshot of it)."""
    import torch
    p = fixture_path(name)
    if not os.path.exists(p):
        return None
    return torch.load(p, map_location="cpu")["generator"]


if __name__ == "__main__":
    for n in ("LJSpeech", "universal"):
        print(make(n, force=True))
