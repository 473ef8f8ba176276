"""Import the UNMODIFIED reference (ming024/FastSpeech2) read-only, for pinning the oracle.  TEST INFRASTRUCTURE.

Only usable where the reference tree exists (this container: /root/reference, or $FS2_REFERENCE).
The GPU box has no such tree: nothing in the `-m gpu` tests, smoke() or bench.py calls this.

The reference's hot-path modules import three front-end/plotting packages that are not installed
(text/cleaners.py:18 unidecode, text/numbers.py:3 inflect, utils/tools.py:7-9 matplotlib); they are
stubbed in sys.modules -- the reference source itself is untouched.  Its modules also bind a global
`device` at import (model/modules.py:14, utils/tools.py:15), so run this on a CUDA-less process.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FS2_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "model")) and os.path.isdir(os.path.join(REFERENCE_ROOT, "hifigan"))


def _stub():
    for name in ("unidecode", "inflect", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["unidecode"].unidecode = lambda s: s
    sys.modules["inflect"].engine = lambda: None
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


_loaded = None


def load():
    """Returns (FastSpeech2 class, hifigan module) from the reference tree."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _stub()
    for m in ("model", "hifigan", "transformer", "utils", "text"):
        if m in sys.modules:
            mod = sys.modules[m]
            where = [getattr(mod, "__file__", None) or ""] + list(getattr(mod, "__path__", []) or [])
            if not any(REFERENCE_ROOT in w for w in where):
                raise RuntimeError(f"module {m!r} already imported from elsewhere; run the reference in its own process")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import hifigan                                   # noqa: E402
    from model import FastSpeech2                    # noqa: E402
    _loaded = (FastSpeech2, hifigan)
    return _loaded
