"""Shared helpers for the parity tests."""
import os

import numpy as np

from golden_io import load_tree

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

_cache = {}


def golden(name):
  if name not in _cache:
    _cache[name] = load_tree(os.path.join(GOLDEN, name + '.npz'))
  return _cache[name]


def relmax(got, want):
  """max|got-want| / max|want|  (per-field max-norm relative error)."""
  got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
  if got.shape != want.shape:
    return np.inf
  if np.array_equal(got, want, equal_nan=True):
    return 0.0
  if not np.array_equal(np.isnan(got), np.isnan(want)):
    return np.inf
  scale = np.nanmax(np.abs(want))
  return float(np.nanmax(np.abs(got - want)) / (scale if scale > 0 else 1.0))
