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


def unit_column(backend, lib_call):
  pass
