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
