"""`from Matcher import Matcher` — put this directory on sys.path (ahead of the reference checkout) and every script
written against the reference picks up the B200 implementation unchanged."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import pymcl_b200 as _pkg  # noqa: E402,F401
from pymcl_b200.Matcher import Matcher, extension  # noqa: E402,F401
