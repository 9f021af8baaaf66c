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
