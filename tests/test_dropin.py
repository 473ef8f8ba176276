"""CPU: the drop-in installer binds the names the reference imports (utils/model.py:7-8) to our modules."""
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_install_binds_model_and_hifigan():
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import fastspeech2_b200.dropin as d; d.install()\n"
        "import hifigan\n"
        "from model import FastSpeech2, ScheduledOptim, FastSpeech2Loss\n"
        "assert FastSpeech2.__module__.startswith('fastspeech2_b200.') and hifigan.Generator.__module__.startswith('fastspeech2_b200.')\n"
        "h = hifigan.AttrDict({'a': 1}); assert h.a == 1\n"
        "try:\n    ScheduledOptim(None, None, None, 0)\nexcept NotImplementedError:\n    print('ok')\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr


def test_install_refuses_to_shadow_an_imported_reference_module():
    code = (
        "import sys, types; sys.path.insert(0, %r)\n"
        "sys.modules['model'] = types.ModuleType('model')\n"
        "import fastspeech2_b200.dropin as d\n"
        "try:\n    d.install()\nexcept RuntimeError:\n    print('refused')\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "refused" in r.stdout, r.stderr


def test_reference_factories_build_our_modules():
    """The reference's own utils.model.get_model / get_vocoder code path (utils/model.py:11-71) constructs, loads and
    prepares OUR modules (needs the reference tree; runs on CPU up to, not including, the forward)."""
    import pytest
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present")
    code = r'''
import sys, os, types, json, zipfile, io, tempfile
sys.path.insert(0, %r)
from oracle import ref_import
ref_import._stub()
import fastspeech2_b200.dropin as d; d.install()
REF = ref_import.REFERENCE_ROOT
sys.path.insert(0, REF)
os.chdir(REF)                                   # the reference uses ./preprocessed_data and hifigan/config.json relative paths
import yaml, torch
from utils.model import get_model            # the reference's factory, untouched
pc = yaml.safe_load(open("config/LJSpeech/preprocess.yaml")); mc = yaml.safe_load(open("config/LJSpeech/model.yaml")); tc = yaml.safe_load(open("config/LJSpeech/train.yaml"))
args = types.SimpleNamespace(restore_step=0)
m = get_model(args, (pc, mc, tc), torch.device("cpu"), train=False)
assert type(m).__module__.startswith("fastspeech2_b200.") and not m.training
# vocoder: replicate utils/model.py:58-69 with the shipped checkpoint (unzipped in memory; the reference reads an unzipped file)
import hifigan
cfg = hifigan.AttrDict(json.load(open("hifigan/config.json")))
voc = hifigan.Generator(cfg)
z = zipfile.ZipFile("hifigan/generator_LJSpeech.pth.tar.zip")
ckpt = torch.load(io.BytesIO(z.read(z.namelist()[0])), map_location="cpu")
voc.load_state_dict(ckpt["generator"]); voc.eval(); voc.remove_weight_norm(); voc.to(torch.device("cpu"))
assert "conv_pre.weight" in voc.state_dict() and "conv_pre.weight_g" not in voc.state_dict()
print("factories ok", sum(p.numel() for p in m.parameters()), sum(p.numel() for p in voc.parameters()))
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "factories ok 35159361 13926017" in r.stdout, r.stdout + r.stderr


def test_patch_text_dataset_rebinds_the_reference_name():
    """`synthesize.py:14` does `from dataset import TextDataset`: after dropin.patch_text_dataset() that name is the batch front-end's
    class, and the stock DataLoader line of synthesize.py:193-198 runs on it."""
    import pytest
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present")
    code = r'''
import sys, os
sys.path.insert(0, %r)
from oracle import ref_import
ref_import._stub()
REF = ref_import.REFERENCE_ROOT
sys.path.insert(0, REF); os.chdir(REF)
import fastspeech2_b200.dropin as d
assert d.patch_text_dataset()
from dataset import TextDataset
assert TextDataset.__module__ == "fastspeech2_b200.frontend"
import yaml
from torch.utils.data import DataLoader
pc = yaml.safe_load(open("config/LJSpeech/preprocess.yaml"))
ds = TextDataset("preprocessed_data/LJSpeech/val.txt", pc)
b = next(iter(DataLoader(ds, batch_size=8, collate_fn=ds.collate_fn)))
assert len(b) == 6 and b[3].shape == (8, b[5])
print("ok")
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
