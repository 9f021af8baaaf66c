"""Parity checks shared by the CPU (warp-emulator) and GPU (C ABI on cuda:0) test files."""
import numpy as np

from helpers import golden, relmax, spec_from_cases
from pymoc_b200.ensemble import Ensemble

# north_star: 1e-10 relative (max-norm per field) on the explicit path
TOL = 1e-10


def run_against_golden(backend, name, nmax, tol=TOL, chunked=False):
  """Run every member of a golden fixture as one ensemble and compare at each checkpoint."""
  tree = golden(name)
  members = list(tree['members'].items())
  ens = Ensemble(spec_from_cases([d['case'] for _, d in members]), backend=backend)
  done, worst = 0, 0.0
  for n in sorted(int(k) for k in members[0][1]['runs']):
    if n > nmax:
      break
    if chunked:  # many short launches: exercises the carried streamfunctions between launches
      while done < n:
        step = min(7, n - done)
        ens.run(step)
        done += step
    else:
      ens.run(n - done)
      done = n
    got = {**ens.state(), **ens.diagnostics()}
    for i, (m, d) in enumerate(members):
      for key, want in d['runs'][str(n)].items():
        if key not in got:
          continue
        err = relmax(got[key][i], want)
        worst = max(worst, err)
        assert err < tol, '%s member %s step %d field %s: rel err %.3e' % (name, m, n, key, err)
    assert not (got['status'] & 1).any(), 'NaN status raised'
  return worst



def diag_and_pickup_files(backend, tmpdir):
  """The --diagfile / --pickup_save_file archives of examples/run_JansenNadeau_2018.py:266-272 for
  the literal script configuration, 360 iterations with Diag_iters = 120, against the files the
  reference writes (golden c4_diags.npz); then a restart from the pickup."""
  import os

  from pymoc_b200 import pickup
  tree = golden('c4_diags')
  ens = Ensemble(spec_from_cases([tree['case']]), backend=backend)
  rec = pickup.DiagRecorder(ens, int(tree['diag_iters'])).run(int(tree['total_iters']))
  dpath, ppath = os.path.join(tmpdir, 'diags.npz'), os.path.join(tmpdir, 'pickup.npz')
  rec.save(dpath, member=0)
  pickup.save_pickup(ens, ppath, member=0)
  got, want = np.load(dpath), tree['files']['diagfile']
  assert sorted(got.files) == sorted(want), (got.files, sorted(want))
  for k in want:
    assert np.shape(got[k]) == np.shape(want[k]), (k, np.shape(got[k]), np.shape(want[k]))
    assert relmax(got[k], want[k]) < TOL, (k, relmax(got[k], want[k]))
  gotp = np.load(ppath)
  for k, v in tree['files']['pickup'].items():
    assert relmax(gotp[k], v) < TOL, k
  # restart: a fresh ensemble picked up from the file continues like the one that wrote it
  # (the scripts restart their iteration counter, so 360 must be -- and is -- a multiple of K)
  again = Ensemble(spec_from_cases([tree['case']]), backend=backend)
  pickup.load_pickup(again, ppath)
  assert again.it == 0
  again.run(24)
  ens.run(24)
  for k, v in ens.state().items():
    assert np.array_equal(v, again.state()[k]), k


def edge_sizes(backend, wide=True):
  """Sizes at the dispatch boundaries of the kernels against the live oracle: levels per lane 2..8 (nz up to
  256) with member counts that do not fill a CTA, the smallest grids, and the block-per-member kernels at
  the boundaries of their levels-per-thread variants."""
  import warnings

  from oracle import pymoc_oracle as O
  from pymoc_b200 import configs
  warnings.filterwarnings('ignore')
  worst = 0.0

  def check(spec, nsteps, members, keys):
    nonlocal worst
    ens = Ensemble(spec, backend=backend)
    ens.run(nsteps)
    got = {**ens.state(), **ens.diagnostics()}
    for m in members:
      want = O.run_coupled(spec.member_case(m), nsteps, O.REFERENCE)
      for key in keys:
        err = relmax(got[key][m], want[key])
        worst = max(worst, err)
        assert err < TOL, (spec.name, spec.nz, m, key, err)

  # column + Psi_SO: every levels-per-lane variant, 3 members (a partly filled CTA), smallest channel grid
  for nz, ny in ((5, 3), (33, 40), (64, 7), (65, 40), (129, 40), (255, 40), (256, 64)):
    with configs.members(0, 3):
      spec = configs.c2_column_so(4, nz=nz, ny=ny, dt_days=0.5)
    check(spec, 2 * spec.K + 3 if spec.K < 40 else 75, (0, 2), ('b_basin', 'Psi_so', 'Psi_Ek', 'Psi_GM'))
  # 'jn' order at the warp kernels' largest grid and at a tiny one
  for nz, dt_days in ((256, 2.), (12, 30.)):
    with configs.members(1, 4):
      spec = configs.c5_single_global_basin(4, nz=nz, dt_days=dt_days, axes=(2, 2, 1, 1), kapfac_max=1.)
    check(spec, 30, (0, 2), ('b_basin', 'b_north', 'bs_ml', 'Psi_iso_b', 'Psi_so'))
  if wide:  # block-per-member kernels: first size, boundaries of the 4 / 8 / 16 levels-per-thread variants
    for nz, dt_days in ((257, 2.), (1024, 0.1), (1025, 0.1), (2049, 0.02)):
      spec = configs.c5_single_global_basin(1, nz=nz, dt_days=dt_days)
      check(spec, 25, (0,), ('b_basin', 'b_north', 'bs_ml', 'Psi_iso_b', 'Psi_so'))
  return worst
