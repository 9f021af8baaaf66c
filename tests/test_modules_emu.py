"""Drop-in classes on the CPU warp emulator of the kernel sources (no GPU needed)."""
import pytest

import module_checks as mc
from emu.emu_backend import EmuBackend
from pymoc_b200.modules import _dispatch


@pytest.fixture(scope='module', autouse=True)
def emu_backend():
  _dispatch._set_backend(EmuBackend())
  yield
  _dispatch._set_backend(None)


def test_column_init():
  mc.column_init_errors()


def test_column_reference_tests():
  mc.column_reference_tests()


def test_column_golden_units():
  mc.column_golden_units()


def test_thermwind():
  mc.thermwind_checks()


def test_so():
  mc.so_checks()


def test_ml():
  mc.ml_checks()


def test_psib_edge_cases():
  mc.psib_edge_cases()


def test_host_side_method_surface():
  mc.host_api_checks()


def test_dropin_example_script_matches_batched_engine():
  """examples/twocol_plusSO_dropin.py (the reference script's loop over the drop-in classes, one launch per
  method call) ends where the fused kernel ends for the same member."""
  import importlib.util
  import os

  import numpy as np
  from emu.emu_backend import EmuBackend
  from pymoc_b200 import configs
  from pymoc_b200.ensemble import Ensemble
  from pymoc_b200.modules import _dispatch
  _dispatch._set_backend(EmuBackend())
  path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'examples', 'twocol_plusSO_dropin.py')
  spec_ = importlib.util.spec_from_file_location('dropin_example', path)
  mod = importlib.util.module_from_spec(spec_)
  spec_.loader.exec_module(mod)
  out = mod.main(49, c=None, nodal=True)
  ens = Ensemble(configs.c3_twocol_so(1), backend=EmuBackend())
  ens.run(49)
  got = {**ens.state(), **ens.diagnostics()}
  for a, b in (('b_basin', 'b_basin'), ('b_north', 'b_north'), ('Psi', 'Psi_tw'), ('Psi_SO', 'Psi_so')):
    err = np.abs(out[a] - got[b][0]).max() / np.abs(got[b][0]).max()
    assert err < 1e-10, (a, err)
