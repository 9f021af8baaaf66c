"""One short workload for ncu: build the ensemble, warm-up launches, one measured launch.
usage: prof_one.py <workload> <members> <steps of the measured launch> [spin-up steps]"""
import sys
sys.path.insert(0, '.')
from pymoc_b200 import configs
from pymoc_b200.ensemble import Ensemble
name, M, nt = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
spin = int(sys.argv[4]) if len(sys.argv) > 4 else nt
mk = {'C2': configs.c2_column_so, 'C3': configs.c3_twocol_so, 'C4': configs.c4_jansen_nadeau, 'C5': configs.c5_single_global_basin,
      'C5_4096': lambda M: configs.c5_single_global_basin(M, nz=4096, dt_days=0.01, kapfac_max=1.),
      'C3_bvp': lambda M: configs.c3_twocol_so(M, c=0.1), 'twobasin': configs.twobasin,
      'C1': configs.c1_timestepping}[name]
spec = mk(M)
ens = Ensemble(spec)
if spec.order == 'post':
  ens.diagnose()
ens.run(spin)   # spin-up / warm-up
ens.run(nt)   # the launch ncu captures (-s skips the ones before it)
print('ok', name, M, nt, int((ens.diagnostics()['status'] & 1).sum()))
