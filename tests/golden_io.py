"""Nested-dict <-> flat ``.npz`` helpers for the golden fixtures (no pickle)."""
import numpy as np

_NONE = '__none__'


def _flatten(tree, prefix, out, nones):
  for key, val in tree.items():
    path = prefix + str(key)
    if isinstance(val, dict):
      _flatten(val, path + '/', out, nones)
    elif val is None:
      nones.append(path)
    else:
      out[path] = np.asarray(val)


def save_tree(path, tree):
  out, nones = {}, []
  _flatten(tree, '', out, nones)
  out[_NONE] = np.array(nones, dtype='U')
  np.savez_compressed(path, **out)


def load_tree(path):
  tree = {}
  with np.load(path, allow_pickle=False) as data:
    nones = [str(s) for s in data[_NONE]] if _NONE in data else []
    items = [(k, data[k]) for k in data.files if k != _NONE] + [(k, None) for k in nones]
  for key, val in items:
    node = tree
    parts = key.split('/')
    for p in parts[:-1]:
      node = node.setdefault(p, {})
    if val is not None and val.ndim == 0:
      val = val.item()
    node[parts[-1]] = val
  return tree
