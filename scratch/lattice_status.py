import sys, time, numpy as np
sys.path.insert(0,'.')
from pymoc_b200 import configs
from pymoc_b200.ensemble import Ensemble
import torch
out = {}
for name, spec, steps in (('C3', configs.c3_twocol_so(32768), (2400, 12000, 12000)), ('C4', configs.c4_jansen_nadeau(32768), (2400, 4800)), ('C5', configs.c5_single_global_basin(4096), (2400, 4800))):
    ens = Ensemble(spec)
    for n in steps:
        torch.cuda.synchronize(); t = time.time()
        ens.run(n)
        torch.cuda.synchronize(); dt = time.time() - t
        st = ens.diagnostics()['status']
        out['%s_status_%d' % (name, ens.it)] = st
        print(name, 'after', ens.it, 'time %.3fs' % dt, 'rate %.3e' % (spec.M * n / dt), 'status bits', {b: int(((st & b) != 0).sum()) for b in (1, 2, 4, 8, 16)}, flush=True)
    for k, v in spec.sweep.items():
        out['%s_%s' % (name, k)] = v
np.savez_compressed('gpurun_out/lattice_status.npz', **out)
