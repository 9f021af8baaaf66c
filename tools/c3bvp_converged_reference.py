import sys, warnings, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
warnings.filterwarnings('ignore')
from scipy import integrate
from emu.emu_backend import EmuBackend
from helpers import relmax, spec_from_cases
from oracle import pymoc_oracle as O
from parity_common import sample_cases
import bench
from pymoc_b200 import configs
from pymoc_b200.ensemble import Ensemble
m = int(sys.argv[1]); n = int(sys.argv[2])
with configs.members(m, m + 1):
  spec = bench.WORKLOADS['C3_bvp'][0](262144)
case = spec.member_case(0)
print({k: float(v[0]) for k, v in spec.sweep.items()})
ens = Ensemble(spec, backend=EmuBackend()); ens.run(n)
got = {**ens.state(), **ens.diagnostics()}
t = time.time(); ref = O.run_coupled(case, n, O.REFERENCE); print('oracle %.0fs' % (time.time() - t))
# the same oracle with solve_bvp run to a tight tolerance in the smoother
orig = integrate.solve_bvp
def tight(fun, bc, x, y, **kw):
  if x.size == case['z'].size and 'tol' not in kw and getattr(tight, 'on', False):
    kw['tol'] = 1e-9; kw['max_nodes'] = 100000
  return orig(fun, bc, x, y, **kw)
# only the smoother's call should be tightened: thermwind's linear problem converges in one iteration either way
tight.on = True
integrate.solve_bvp = tight
t = time.time(); conv = O.run_coupled(case, n, O.REFERENCE); print('tight oracle %.0fs' % (time.time() - t))
integrate.solve_bvp = orig
for k in ('b_basin', 'b_north', 'Psi_so', 'Psi_iso_b'):
  print(k, 'kernel vs reference(tol=1e-3) %.2e | kernel vs reference algorithm with solve_bvp(tol=1e-9) %.2e | the two references %.2e' % (
    relmax(got[k][0], ref[k]), relmax(got[k][0], conv[k]), relmax(ref[k], conv[k])))
