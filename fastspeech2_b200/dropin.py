"""Make the reference's `import model` / `import hifigan` resolve to the B200-native drop-ins.

    import fastspeech2_b200.dropin as dropin; dropin.install()           # programmatic
    python -m fastspeech2_b200.dropin synthesize.py --text ... -p ... -m ... -t ...   # run the untouched reference CLI

`utils/model.py:7-8` does `import hifigan` and `from model import FastSpeech2, ScheduledOptim`; both names are bound here
to `fastspeech2_b200.hifigan` / `fastspeech2_b200.model`, whose classes keep the reference's constructor arguments,
state_dict keys, forward signatures and return values.
"""
from __future__ import annotations

import os
import runpy
import sys


def install(force: bool = False) -> None:
    import fastspeech2_b200.hifigan as b_hifigan
    import fastspeech2_b200.model as b_model
    for name, mod in (("model", b_model), ("hifigan", b_hifigan)):
        cur = sys.modules.get(name)
        if cur is not None and cur is not mod and not force:
            raise RuntimeError(f"module {name!r} is already imported from {getattr(cur, '__file__', '?')}; call install() first")
        sys.modules[name] = mod


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m fastspeech2_b200.dropin <reference script, e.g. synthesize.py> [script args...]")
    script = os.path.abspath(argv[0])
    install()
    sys.argv = [script] + argv[1:]
    sys.path.insert(0, os.path.dirname(script))       # the reference resolves `utils`, `text`, `dataset` relative to itself
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
