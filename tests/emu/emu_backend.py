"""Test-only backend that runs the kernel sources through the CPU warp emulator."""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build_emu  # noqa: E402

from pymoc_b200 import _abi  # noqa: E402

_lib = None


def emu_lib():
  global _lib
  if _lib is None:
    _lib = _abi.declare(ctypes.CDLL(build_emu.build()))
  return _lib


class EmuBackend:
  def __init__(self):
    self.lib = emu_lib()

  def upload(self, arr):
    return np.array(arr, copy=True, order='C')

  def zeros(self, shape, dtype=np.float64):
    return np.zeros(shape, dtype=dtype)

  @staticmethod
  def ptr(buf):
    return None if buf is None else buf.ctypes.data

  def download(self, buf):
    return buf.copy()

  def assign(self, buf, arr):
    buf[...] = arr

  def stream(self):
    return None

  def sync(self):
    pass

  def guard(self):
    import contextlib
    return contextlib.nullcontext()
