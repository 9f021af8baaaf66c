"""GPU parity tests proper: the C ABI of libpymoc_b200.so on cuda:0 vs the reference.

Golden fixtures come from the unmodified reference (tests/golden/make_golden.py); the live
oracle (oracle/pymoc_oracle.py, pinned to the same fixtures) covers members the fixtures
do not hold.  Tolerance: 1e-10 relative, max-norm per field (north_star, explicit path).
"""
import ctypes
import warnings

import numpy as np
import pytest

from helpers import relmax
from parity_common import TOL, diag_and_pickup_files, edge_sizes, run_against_golden

pytestmark = pytest.mark.gpu
warnings.filterwarnings('ignore', category=RuntimeWarning)


@pytest.fixture(scope='module')
def cuda():
  from pymoc_b200.backend import CudaBackend
  return CudaBackend()


@pytest.mark.parametrize('name,nmax', [('c1', 1200), ('c2', 2160), ('twocol', 480), ('c3', 2400), ('c4', 2400),
                                       ('c4_literal', 1200), ('c5', 480), ('c5_wide', 1000), ('c5_4096', 40), ('c5_4096_k20', 45),
                                       ('twobasin', 1200)])
def test_fused_kernel_vs_reference(cuda, name, nmax):
  worst = run_against_golden(cuda, name, nmax)
  print('%s: worst relative error %.2e' % (name, worst))


def test_f2010_smoother_literal_c3(cuda):
  """examples/example_twocol_plusSO.py as written (c=0.1, bvp_with_Ek=True): the reference's adaptive
  solve_bvp (tol=1e-3) against the converged solution of the same ODE; stated tolerance 1e-5."""
  worst = run_against_golden(cuda, 'c3_bvp', 240, tol=1e-5)
  print('c3_bvp: worst relative error %.2e' % worst)


@pytest.mark.parametrize('name,nmax', [('c2', 73), ('c3', 25), ('c4', 600)])
def test_short_launches_carry_streamfunctions(cuda, name, nmax):
  run_against_golden(cuda, name, nmax, chunked=True)


def test_ensemble_members_vs_live_oracle(cuda):
  """A 256-member tau x kappa x bs_north x A lattice of the C3 model: sampled members against
  the oracle, and the whole ensemble for finiteness and lattice symmetry."""
  from oracle import pymoc_oracle as O
  from pymoc_b200 import configs
  from pymoc_b200.ensemble import Ensemble
  spec = configs.c3_twocol_so(256, axes=(4, 4, 4, 4))
  ens = Ensemble(spec, backend=cuda)
  ens.run(121)
  got = {**ens.state(), **ens.diagnostics()}
  assert np.isfinite(got['b_basin']).all() and np.isfinite(got['Psi_iso_b']).all()
  for m in (0, 37, 101, 200, 255):
    want = O.run_coupled(spec.member_case(m), 121, O.REFERENCE)
    for key in ('b_basin', 'b_north', 'Psi_tw', 'Psi_iso_b', 'Psi_iso_n', 'Psi_so', 'psib'):
      assert relmax(got[key][m], want[key]) < TOL, (m, key, relmax(got[key][m], want[key]))


def test_full_size_properties(cuda):
  """BASELINE size (65,536 members, nz=200): size-independent properties.
  (1) duplicated parameters give bit-identical members; (2) splitting a run into launches
  changes nothing bit-wise; (3) the surface/bottom boundary values are the exact copies the
  reference makes (b[-1] == bs, b[0] == bbot); (4) Psi_so[0] == 0."""
  from pymoc_b200 import configs
  from pymoc_b200.ensemble import Ensemble
  spec = configs.c2_column_so(65536)
  a = Ensemble(spec, backend=cuda)
  a.run(145)
  b = Ensemble(spec, backend=cuda)
  b.run(72)
  b.run(1)
  b.run(72)
  sa, sb = a.state()['b_basin'], b.state()['b_basin']
  assert np.array_equal(sa, sb)
  assert np.array_equal(a.diagnostics()['Psi_so'], b.diagnostics()['Psi_so'])
  assert np.all(sa[:, -1] == 0.03) and np.all(sa[:, 0] == 0.0)
  assert np.all(a.diagnostics()['Psi_so'][:, 0] == 0.0)
  # members 0..255 share tau (first lattice axis) but not kappa: all distinct
  assert len({sa[i].tobytes() for i in range(256)}) == 256
  assert np.isfinite(sa).all()


def test_host_buffer_entry_point(cuda):
  """pmoc_model_run_host (host pointers, copies inside) == device path, bit for bit."""
  from emu.emu_backend import EmuBackend  # only for its numpy buffer handling
  from pymoc_b200 import _lib, configs
  from pymoc_b200.ensemble import Ensemble

  class HostBuffers(EmuBackend):
    def __init__(self, lib):
      self.lib = lib

  # 16 members: one block; 32,768 members: four blocks pipelined over three streams
  for spec, keys in ((configs.c3_twocol_so(16, axes=(2, 2, 2, 2)), ('Psi_tw', 'Psi_iso_b', 'Psi_so', 'psib')),
                     (configs.c2_column_so(32768), ('Psi_so', 'Psi_Ek', 'Psi_GM')),
                     (configs.c4_jansen_nadeau(16384), ('Psi_iso_b', 'Psi_so', 'Psi_s', 'bbot_basin')),
                     # block-per-member kernels: the library allocates their scratch per stream (two blocks)
                     (configs.c5_single_global_basin(8192, nz=320, dt_days=1., kapfac_max=1.), ('Psi_iso_b', 'Psi_so'))):
    dev = Ensemble(spec, backend=cuda)
    dev.run(30)
    host = Ensemble(spec, backend=HostBuffers(cuda.lib))
    _lib.check(cuda.lib.pmoc_model_run_host(ctypes.byref(host.model), 0, 13))
    _lib.check(cuda.lib.pmoc_model_run_host(ctypes.byref(host.model), 13, 17))
    for key, val in dev.state().items():
      assert np.array_equal(val, host.state()[key], equal_nan=True), key
    for key in keys:
      assert np.array_equal(dev.diagnostics()[key], host.diagnostics()[key], equal_nan=True), key


def test_diagnostics_and_pickup_wire_format(cuda, tmp_path):
  diag_and_pickup_files(cuda, str(tmp_path))


def test_full_size_c3_properties(cuda):
  """BASELINE configs[2] size (262,144 members of the two-column + Psi_SO model): (1) splitting a run into
  launches changes nothing bit-wise (the carried streamfunctions are exact copies); (2) members that differ
  only in a parameter the northern column does not see ... all members distinct, all finite; (3) boundary
  values are the exact copies the reference makes; (4) every sampled member matches the live oracle."""
  from oracle import pymoc_oracle as O
  from pymoc_b200 import configs
  from pymoc_b200.ensemble import Ensemble
  M = 262144
  spec = configs.c3_twocol_so(M)
  a = Ensemble(spec, backend=cuda)
  a.run(49)
  b = Ensemble(spec, backend=cuda)
  for n in (24, 1, 24):
    b.run(n)
  sa, sb = a.state(), b.state()
  for k in sa:
    assert np.array_equal(sa[k], sb[k]), k
  da, db = a.diagnostics(), b.diagnostics()
  for k in ('Psi_tw', 'Psi_iso_b', 'Psi_iso_n', 'Psi_so', 'psib'):
    assert np.array_equal(da[k], db[k]), k
  assert not (da['status'] & 1).any()
  assert np.isfinite(sa['b_basin']).all() and np.isfinite(sa['b_north']).all()
  assert np.all(sa['b_basin'][:, -1] == 0.03) and np.all(sa['b_basin'][:, 0] == 0.0)
  assert np.all(da['Psi_so'][:, 0] == 0.0) and np.all(da['Psi_tw'][:, 0] == 0.0)
  for m in (0, 12345, 99999, M - 1):
    want = O.run_coupled(spec.member_case(m), 49, O.REFERENCE)
    for key in ('b_basin', 'b_north', 'Psi_iso_b', 'Psi_so'):
      got = sa[key][m] if key in sa else da[key][m]
      assert relmax(got, want[key]) < TOL, (m, key, relmax(got, want[key]))


def test_jn_and_wide_split_launches_bitwise(cuda):
  """'jn' order (C4, 65,536 members) and the block-per-member kernels (nz = 320): a run split into
  launches at and between diagnosis iterations equals the single launch bit for bit, including the
  carried bottom-boundary switches (bbot, kappa variant) and the mixed layer."""
  from pymoc_b200 import configs
  from pymoc_b200.ensemble import Ensemble
  for spec, n, cuts in ((configs.c4_jansen_nadeau(65536), 61, (12, 7, 30, 12)),
                        (configs.c5_single_global_basin(128, nz=320, dt_days=1., kapfac_max=1.), 1450, (720, 5, 725))):
    a = Ensemble(spec, backend=cuda)
    a.run(n)
    b = Ensemble(spec, backend=cuda)
    assert sum(cuts) == n
    for c in cuts:
      b.run(c)
    sa, sb = a.state(), b.state()
    for k in sa:
      assert np.array_equal(sa[k], sb[k], equal_nan=True), (spec.name, k)
    da, db = a.diagnostics(), b.diagnostics()
    for k in ('Psi_iso_b', 'Psi_iso_n', 'Psi_so', 'bbot_basin', 'bbot_north'):
      assert np.array_equal(da[k], db[k], equal_nan=True), (spec.name, k)


def test_edge_sizes_vs_live_oracle(cuda):
  print('edge sizes: worst relative error %.2e' % edge_sizes(cuda, wide=True))


# Bench-lattice samples: seeded-random members of the very lattices bench.py steps,
# every one against the live oracle; see parity_common.lattice_sample for the contract.
@pytest.mark.parametrize('workload,M,nsample,nsteps,tol,max_flagged', [
    ('C1', 65536, 64, 600, TOL, 0.02),
    ('C2', 65536, 256, 600, TOL, 0.02),
    ('C3', 262144, 256, 600, TOL, 0.02),
    ('twobasin', 32768, 256, 600, TOL, 0.02),
    ('C4', 32768, 256, 600, TOL, 0.10),
    ('C5', 32768, 256, 600, TOL, 0.02),
    ('C3_bvp', 262144, 32, 2400, 1e-5, 0.02),  # the literal script (F2010 smoother): stated tolerance, 100 refreshes
])
def test_bench_lattice_sample_vs_live_oracle(cuda, workload, M, nsample, nsteps, tol, max_flagged):
  from parity_common import lattice_sample
  rep = lattice_sample(cuda, workload, M, nsample, nsteps, seed=20261018, tol=tol, max_flagged=max_flagged)
  print('lattice sample %s' % rep)


def test_c4_lattice_long_run(cuda):
  """C4 at the bench's own length (2 400 steps: steady state, where the no-flux bottom ties develop)."""
  from parity_common import lattice_sample
  rep = lattice_sample(cuda, 'C4', 32768, 64, 2400, seed=4, max_flagged=0.15)
  print('lattice sample %s' % rep)


def test_f2010_smoother_is_the_converged_solution(cuda):
  from parity_common import f2010_smoother_converged
  print('F2010 smoother vs solve_bvp(tol=1e-8): worst %.2e' % f2010_smoother_converged(cuda, bvp_tol=1e-8, tol=1e-8))


def test_host_handle_matches_device_path(cuda):
  """pmoc_host_open / pmoc_host_step / pmoc_host_close (persistent host-buffer handle: parameters uploaded once,
  per call only the requested array classes move) == the device path, bit for bit -- with nothing pulled between
  calls, with state + streamfunctions pulled and pushed back every call, and across a diagnosis boundary."""
  from pymoc_b200 import configs
  from pymoc_b200.ensemble import Ensemble, HostEnsemble
  S, P, D = HostEnsemble.IO_STATE, HostEnsemble.IO_PSI, HostEnsemble.IO_DIAG
  for spec, keys in ((configs.c3_twocol_so(16, axes=(2, 2, 2, 2)), ('Psi_tw', 'Psi_iso_b', 'Psi_so', 'psib')),
                     (configs.c2_column_so(32768), ('Psi_so', 'Psi_Ek', 'Psi_GM')),
                     (configs.c4_jansen_nadeau(16384), ('Psi_iso_b', 'Psi_so', 'Psi_s', 'bbot_basin')),
                     (configs.c5_single_global_basin(8192, nz=320, dt_days=1., kapfac_max=1.), ('Psi_iso_b', 'Psi_so'))):
    dev = Ensemble(spec, backend=cuda)
    dev.run(30)
    a = HostEnsemble(spec)
    a.run(13, pull=0)
    assert a.last_bytes() == (0, 0)  # nothing moved: parameters and state are resident
    a.run(17, pull=S | P | D)
    b = HostEnsemble(spec)
    b.run(13, pull=S | P)
    b.run(17, push=S | P, pull=S | P | D)  # the host copy is the truth: up, 17 steps, down
    h2d, d2h = b.last_bytes()
    assert h2d > 0 and d2h > h2d
    for ens in (a, b):
      for key, val in dev.state().items():
        assert np.array_equal(val, ens.state()[key], equal_nan=True), (spec.name, key)
      for key in keys:
        assert np.array_equal(dev.diagnostics()[key], ens.diagnostics()[key], equal_nan=True), (spec.name, key)
      assert np.array_equal(dev.diagnostics()['status'], ens.diagnostics()['status'])
      ens.close()


def test_twcol_kernel_sizes_vs_live_oracle(cuda):
  from parity_common import twcol_sizes
  print('twcol sizes: worst relative error %.2e' % twcol_sizes(cuda))


def test_c5_4096_lattice_sample_across_diagnoses(cuda):
  """BASELINE configs[4] (nz = 4096, block-per-member kernels): members of the 16,384-member bench lattice, with the
  streamfunctions re-diagnosed every 20 iterations so that 45 steps cross three diagnoses, against the live oracle."""
  from parity_common import lattice_sample
  rep = lattice_sample(cuda, 'C5_4096', 16384, 6, 45, seed=5, K=20, max_flagged=0.2)
  print('lattice sample %s' % rep)


def test_wide_kernel_serial_schedule_is_bitwise_the_pipelined_one(cuda, monkeypatch):
  """Block-per-member step kernel (nz = 320 and 1 100: four and eight levels per thread, SO_ML on its own warp one step
  behind the columns): the serial fall-back taken when the kappa choice hangs on bs[0], forced through the library's
  test hook, gives the bits of the pipelined schedule."""
  from pymoc_b200 import configs
  from pymoc_b200.ensemble import Ensemble
  for nz, dt_days, n in ((320, 1., 730), (1100, 0.1, 300)):
    spec = configs.c5_single_global_basin(64, nz=nz, dt_days=dt_days, kapfac_max=1.)
    monkeypatch.delenv('PMOC_WIDE_FORCE_LATE', raising=False)
    a = Ensemble(spec, backend=cuda)
    a.run(n)
    monkeypatch.setenv('PMOC_WIDE_FORCE_LATE', '1')
    b = Ensemble(spec, backend=cuda)
    b.run(n)
    for k, v in a.state().items():
      assert np.array_equal(v, b.state()[k], equal_nan=True), (nz, k)
    for k in ('Psi_iso_b', 'Psi_so', 'bbot_basin', 'Psi_s'):
      assert np.array_equal(a.diagnostics()[k], b.diagnostics()[k], equal_nan=True), (nz, k)
