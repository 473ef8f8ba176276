"""Mirror of the reference's vendored `hifigan` package surface (hifigan/__init__.py): Generator and AttrDict."""
from .models import Generator


class AttrDict(dict):
    """dict whose keys are also attributes (hifigan/__init__.py:4-7)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


__all__ = ["Generator", "AttrDict"]
