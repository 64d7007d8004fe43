"""`from analyze import analyzer` — see dropin/Matcher.py."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import pymcl_b200 as _pkg  # noqa: E402,F401
from pymcl_b200.analyze import analyzer, extension  # noqa: E402,F401
