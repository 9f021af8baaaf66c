"""Drop-in classes on cuda:0 through the per-module C-ABI entry points (the reference's own
unit tests + the bit-level fixtures of tests/golden/units.npz)."""
import pytest

import module_checks as mc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', autouse=True)
def cuda_backend():
  from pymoc_b200.backend import CudaBackend
  from pymoc_b200.modules import _dispatch
  _dispatch._set_backend(CudaBackend())
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
