"""b200-mcl-match: B200-native (sm_100a) descriptor-matching hot path of zhangxingshuo/py-mcl.

``Matcher`` / ``analyzer`` keep the reference's API (Matcher.py:35, analyze.py:32); the work is
done by hand-written CUDA kernels in ``csrc/`` behind the C-ABI declared in ``include/mcl.h``
(``libmcl.so``), called through ctypes.  There is no CPU fallback: importing the kernels on a box
without the built library or without an sm_100 GPU raises.
"""
__all__ = ["synth"]
