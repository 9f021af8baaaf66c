"""The C-ABI library loads on a GPU-less box and exports every symbol include/pymoc_b200.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
  from pymoc_b200 import build, _lib
  build.build(verbose=False)
  return _lib.lib()


def test_header_symbols_exported(lib):
  from pymoc_b200 import _abi
  header = open(os.path.join(ROOT, 'include', 'pymoc_b200.h')).read()
  declared = set(re.findall(r'^(?:int|void|uint64_t|const char\*)\s+(pmoc_[a-z0-9_]+)\s*\(', header, flags=re.M))
  assert declared == set(_abi.EXPORTS), declared ^ set(_abi.EXPORTS)
  for name in declared:
    assert hasattr(lib, name), name
  assert lib.pmoc_abi_version() == _abi.ABI_VERSION


def test_struct_layout_matches_header(lib):
  """sizeof(pmoc_model) etc. as the C compiler sees them == the ctypes mirror."""
  import subprocess
  import tempfile
  from pymoc_b200 import _abi
  src = '#include <stdio.h>\n#include "pymoc_b200.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(pmoc_vec), sizeof(pmoc_column), sizeof(pmoc_model));return 0;}\n'
  with tempfile.TemporaryDirectory() as d:
    open(os.path.join(d, 't.c'), 'w').write(src)
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), os.path.join(d, 't.c'), '-o', os.path.join(d, 't')])
    sizes = [int(x) for x in subprocess.check_output([os.path.join(d, 't')]).split()]
  assert sizes == [ctypes.sizeof(_abi.Vec), ctypes.sizeof(_abi.Column), ctypes.sizeof(_abi.Model)]


def test_invalid_arguments_are_rejected_without_a_gpu(lib):
  from pymoc_b200 import _abi
  m = _abi.Model()
  assert lib.pmoc_model_run(ctypes.byref(m), 0, 1, None) == _abi.EINVAL
  assert b'M' in lib.pmoc_last_error()


def test_no_cpu_fallback():
  """Without a GPU the product front end refuses to run (the oracle is never a fallback)."""
  import torch
  if torch.cuda.is_available():
    pytest.skip('GPU present')
  from pymoc_b200 import configs
  from pymoc_b200.ensemble import Ensemble
  with pytest.raises(RuntimeError, match='no CUDA device'):
    Ensemble(configs.c1_timestepping(1))


def test_spec_rejects_arrays_that_are_neither_shared_nor_per_member():
  """a per-member array whose leading dimension is neither 1 nor M would be read out of bounds
  on the device; ModelSpec refuses it (e.g. tau on the ny-point grid passed 1-D, kappa variants passed 2-D)."""
  import numpy as np
  import pytest

  from pymoc_b200 import configs
  from pymoc_b200.spec import ChannelSpec, ColumnSpec, ModelSpec
  good = configs.c2_column_so(8, ntau=4)
  z, y = good.z, good.so.y
  col = ColumnSpec.build(z, np.full((8, z.size), 2e-5), 6e13, 0.03, 0.03 * np.exp(z / 300.))
  with pytest.raises(ValueError, match='so.tau'):  # tau on y as a bare [ny] array: ny != M
    ModelSpec(M=8, z=z, dt=good.dt, K=good.K, basin=col, so=ChannelSpec.build(y, good.so.bs[:1], np.linspace(.1, .2, y.size)))
  with pytest.raises(ValueError, match='basin.kappa'):  # [nvar=2, nz] read as [M=2, nz] but M == 8
    ModelSpec(M=8, z=z, dt=good.dt, K=good.K, basin=ColumnSpec.build(z, np.full((2, z.size), 2e-5), 6e13, 0.03, 0.0 * z),
              so=good.so)
  with pytest.raises(ValueError, match='basin.bs'):
    ModelSpec(M=8, z=z, dt=good.dt, K=good.K, basin=ColumnSpec.build(z, 2e-5, 6e13, np.full(3, 0.03), 0.0 * z), so=good.so)
  # the shared / per-member forms pass
  ModelSpec(M=8, z=z, dt=good.dt, K=good.K, basin=col,
            so=ChannelSpec.build(y, good.so.bs[:1], np.linspace(.1, .2, y.size)[None, :]))
