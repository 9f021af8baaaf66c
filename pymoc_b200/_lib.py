"""Loader of the CUDA library.  There is no CPU fallback: a missing library or GPU raises."""
import ctypes
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('PMOC_B200_LIB') or os.path.join(_HERE, 'libpymoc_b200.so')  # (the override is for tuning builds)
_lib = None


def lib():
  """The loaded ``libpymoc_b200.so`` (hand-written sm_100a kernels behind include/pymoc_b200.h)."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise RuntimeError('pymoc_b200: %s is missing -- run `python -m pymoc_b200.build` (needs nvcc). '
                         'There is no CPU fallback.' % LIB_PATH)
    _lib = _abi.declare(ctypes.CDLL(LIB_PATH))
  return _lib


def check(rc):
  _abi.check(lib(), rc)
