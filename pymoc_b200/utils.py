"""Input normalisation helpers of the reference front end.

Mirrors ``src/pymoc/utils/make_func.py:30-45``, ``src/pymoc/utils/make_array.py:30-37``
and ``src/pymoc/utils/check_numpy_version.py:16-21`` of the reference: the accepted
types (callable / ``np.ndarray`` / ``float`` -- *not* ``int``), the aliasing rule
(an ndarray is handed back as-is, no copy) and the two-element ``TypeError`` are
part of the drop-in contract (``tests/utils/test_make_func.py:9-26``).

Everything here runs on the host at set-up time; per ``north_star`` user callables
never reach the GPU, they are sampled on the grid and shipped as arrays.
"""
import numpy as np

_ERR = 'needs to be either function, numpy array, or float'


def make_func(myst, axis, name):
  """Return ``myst`` as a callable of position along ``axis``."""
  if callable(myst):
    return myst
  if isinstance(myst, np.ndarray):
    return lambda x: np.interp(x, axis, myst)
  if isinstance(myst, float):
    return lambda x: myst + 0 * x
  raise TypeError(name, _ERR)


def make_array(myst, axis, name):
  """Return ``myst`` sampled on ``axis`` (ndarrays are returned un-copied)."""
  if isinstance(myst, np.ndarray):
    return myst
  if callable(myst):
    return myst(axis)
  if isinstance(myst, float):
    return myst + 0 * axis
  raise TypeError(name, _ERR)


def check_numpy_version():
  """True when ``np.gradient`` accepts coordinate arrays (numpy >= 1.13)."""
  major, minor = (int(p) for p in np.version.version.split('.')[:2])
  return not (major <= 1 and minor < 13)
