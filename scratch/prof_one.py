"""One short workload for ncu: build the ensemble, one warm-up launch, one measured launch."""
import sys
sys.path.insert(0, '.')
from pymoc_b200 import configs
from pymoc_b200.ensemble import Ensemble
name, M, nt = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
mk = {'C2': configs.c2_column_so, 'C3': configs.c3_twocol_so, 'C4': configs.c4_jansen_nadeau, 'C5': configs.c5_single_global_basin, 'C5_4096': lambda M: configs.c5_single_global_basin(M, nz=4096, dt_days=0.01, kapfac_max=1.),
      'C1': configs.c1_timestepping}[name]
spec = mk(M)
ens = Ensemble(spec)
if spec.order == 'post':
  ens.diagnose()
ens.run(nt)   # warm-up
ens.run(nt)   # the launch ncu captures (-s skips the ones before it)
print('ok', name, M, nt, int((ens.diagnostics()['status'] & 1).sum()))
