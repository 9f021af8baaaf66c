"""Build ``libpymoc_b200.so`` (hand-written sm_100a CUDA behind a C ABI) in-tree.

``python -m pymoc_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles without a
GPU.  The fused kernel is instantiated once per levels-per-lane value, each in its own
translation unit, and the units are compiled in parallel.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(CSRC, 'build')
LIB = os.path.join(HERE, 'libpymoc_b200.so')
LPLS = (2, 3, 4, 5, 6, 7, 8)
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-fmad=false', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def units():
  """(source, object name, extra defines)"""
  out = [('pmoc_ops.cu', 'ops.o', []), ('pmoc_host.cu', 'host.o', []), ('pmoc_wide.cu', 'wide.o', []),
         ('pmoc_twcol.cu', 'twcol.o', [])]
  out += [('pmoc_model.cu', 'model_%d.o' % n, ['-DPM_LPL=%d' % n]) for n in LPLS]
  return out


def _digest():
  h = hashlib.sha256()
  for name in sorted(os.listdir(CSRC)):
    if name.endswith(('.cu', '.cuh')):
      h.update(open(os.path.join(CSRC, name), 'rb').read())
  h.update(open(os.path.join(os.path.dirname(HERE), 'include', 'pymoc_b200.h'), 'rb').read())
  h.update(' '.join(NVCC_FLAGS).encode())
  return h.hexdigest()


def _compile(src, obj, defs):
  cmd = ['nvcc'] + NVCC_FLAGS + defs + ['-c', os.path.join(CSRC, src), '-o', os.path.join(OBJ, obj)]
  r = subprocess.run(cmd, capture_output=True, text=True)
  open(os.path.join(OBJ, obj + '.log'), 'w').write(r.stdout + r.stderr)
  if r.returncode != 0:
    raise RuntimeError('nvcc failed for %s:\n%s' % (src, (r.stdout + r.stderr)[-4000:]))
  return obj


def build_variant(out, extra_flags, only_lpl=None):
  """Tuning builds: the same sources with extra nvcc flags into another library (loaded with PMOC_B200_LIB=<path>)."""
  global OBJ, LIB, NVCC_FLAGS, LPLS
  saved = OBJ, LIB, NVCC_FLAGS, LPLS
  OBJ, LIB, NVCC_FLAGS = OBJ + '_' + os.path.basename(out), out, NVCC_FLAGS + list(extra_flags)
  try:
    return build(force=False, verbose=True)
  finally:
    OBJ, LIB, NVCC_FLAGS, LPLS = saved


def build(force=False, verbose=True):
  os.makedirs(OBJ, exist_ok=True)
  stamp = os.path.join(OBJ, 'stamp')
  digest = _digest()
  if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
    if verbose:
      print('pymoc_b200: %s is up to date' % LIB)
    return LIB
  with cf.ThreadPoolExecutor(max_workers=min(len(units()), os.cpu_count() or 1)) as ex:
    objs = list(ex.map(lambda u: _compile(*u), units()))
  cmd = ['nvcc', '-shared', '-o', LIB] + [os.path.join(OBJ, o) for o in objs]
  r = subprocess.run(cmd, capture_output=True, text=True)
  if r.returncode != 0:
    raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
  open(stamp, 'w').write(digest)
  if verbose:
    print('pymoc_b200: built %s' % LIB)
  return LIB


if __name__ == '__main__':
  if '--variant' in sys.argv:  # python -m pymoc_b200.build --variant <out.so> <nvcc flag> ...
    i = sys.argv.index('--variant')
    build_variant(sys.argv[i + 1], sys.argv[i + 2:])
  else:
    build(force='--force' in sys.argv)
