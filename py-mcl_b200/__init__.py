"""b200-mcl-match: B200-native (sm_100a) descriptor-matching hot path of zhangxingshuo/py-mcl.

``Matcher`` / ``analyzer`` keep the reference's API (reference Matcher.py:35, analyze.py:32); the work is
done by hand-written CUDA kernels in ``csrc/`` behind the C-ABI declared in ``include/mcl.h``
(``libmcl.so``), called through ctypes (``_lib``).  There is no CPU fallback: using the kernels on a box
without the built library or without an sm_100 GPU raises ``MclError``.

    from pymcl_b200.Matcher import Matcher            # drop-in for `from Matcher import Matcher`
    from pymcl_b200.analyze import analyzer           # drop-in for `from analyze import analyzer`
"""
from . import _lib  # noqa: F401  (does not touch the GPU or load the library at import time)
from ._lib import MclError  # noqa: F401

__all__ = ["Matcher", "analyze", "MclError", "synth", "engine", "build"]


def __getattr__(name):
    # submodules are imported on first use (cv2 is only needed by the reference-facing classes):
    #   from pymcl_b200.Matcher import Matcher      # drop-in for `from Matcher import Matcher`
    #   from pymcl_b200.analyze import analyzer     # drop-in for `from analyze import analyzer`
    if name in ("Matcher", "analyze", "synth", "engine", "build"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
