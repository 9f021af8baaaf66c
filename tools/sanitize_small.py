"""Small launches of every kernel family for compute-sanitizer (racecheck / memcheck / synccheck / initcheck):
C3 (fused warp kernel, 'post' order, remap scratch overlays), C3_bvp (F2010 smoother), C4 ('jn' order: SO_ML,
shared-memory atomics in the remap, scratch overlaid on the column tables), two-basin, C1, and the
block-per-member kernels at nz = 320 (k_wide_refresh / k_wide_steps2: block barriers, shared-memory atomics).
usage: compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from pymoc_b200 import configs  # noqa: E402
from pymoc_b200.ensemble import Ensemble, HostEnsemble  # noqa: E402

cases = [('C3', configs.c3_twocol_so(64, axes=(4, 4, 2, 2)), 49), ('C3_bvp', configs.c3_twocol_so(64, c=0.1, axes=(4, 4, 2, 2)), 49),
         ('C4', configs.c4_jansen_nadeau(64, axes=(2, 2, 2, 2, 4)), 37), ('twobasin', configs.twobasin(32, axes=(4, 4, 2)), 49),
         ('C1', configs.c1_timestepping(64), 20), ('C2', configs.c2_column_so(64), 80),
         ('C5_320', configs.c5_single_global_basin(8, nz=320, dt_days=1., kapfac_max=1., axes=(2, 2, 2, 1)), 730)]
for name, spec, n in cases:
  ens = Ensemble(spec)
  ens.run(n)
  st = ens.state()
  assert all(np.isfinite(v).all() for v in st.values()), name
  print(name, 'ok', {k: int(v) for k, v in zip(*np.unique(ens.diagnostics()['status'], return_counts=True))}, flush=True)
h = HostEnsemble(cases[0][1])
h.run(25)
h.close()
print('host handle ok')
