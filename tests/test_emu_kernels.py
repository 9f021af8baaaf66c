"""Kernel-logic tests on the GPU-less build box (CPU only).

The product kernel sources are compiled with g++ against the warp emulator in tests/emu/
(test infrastructure, not a fallback) and driven through the same C ABI and the same
Python front end as on the GPU; results are compared with the golden fixtures produced by
the reference.  The GPU parity tests proper are in test_gpu_parity.py.
"""
import numpy as np
import pytest

from emu.emu_backend import EmuBackend
from parity_common import diag_and_pickup_files, edge_sizes, run_against_golden


@pytest.fixture(scope='module')
def emu():
  return EmuBackend()


@pytest.mark.parametrize('name,nmax', [('c1', 1200), ('c2', 720), ('twocol', 480), ('c3', 480), ('c4', 600), ('c4_literal', 120),
                                       ('c5', 480), ('c5_wide', 1000), ('c5_4096', 40), ('twobasin', 480)])
def test_fused_kernel_vs_reference(emu, name, nmax):
  run_against_golden(emu, name, nmax)


@pytest.mark.parametrize('name,nmax', [('c2', 73), ('c3', 25), ('c4', 13)])
def test_short_launches_carry_streamfunctions(emu, name, nmax):
  run_against_golden(emu, name, nmax, chunked=True)


def test_f2010_smoother_literal_c3(emu):
  """examples/example_twocol_plusSO.py as written (c=0.1, bvp_with_Ek=True): the reference's adaptive
  solve_bvp (tol=1e-3) against the converged solution of the same ODE; stated tolerance 1e-5."""
  worst = run_against_golden(emu, 'c3_bvp', 240, tol=1e-5)
  assert worst > 1e-12  # the two are different discretisations; a tiny number would mean the test is vacuous


def test_diagnostics_and_pickup_wire_format(emu, tmp_path):
  diag_and_pickup_files(emu, str(tmp_path))


def test_jn_split_launches_bitwise(emu):
  """'jn' order: a run split into launches at and between diagnosis iterations equals the single launch."""
  import numpy as np
  from pymoc_b200 import configs
  from pymoc_b200.ensemble import Ensemble
  spec = configs.c4_jansen_nadeau(4, axes=(2, 2, 1, 1, 1))
  a = Ensemble(spec, backend=emu)
  a.run(37)
  b = Ensemble(spec, backend=emu)
  for c in (12, 7, 18):
    b.run(c)
  for k, v in a.state().items():
    assert np.array_equal(v, b.state()[k], equal_nan=True), k
  for k in ('Psi_iso_b', 'Psi_so', 'bbot_basin', 'bbot_north'):
    assert np.array_equal(a.diagnostics()[k], b.diagnostics()[k], equal_nan=True), k


def test_edge_sizes_vs_live_oracle(emu):
  edge_sizes(emu, wide=True)
