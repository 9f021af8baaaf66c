import sys, time, numpy as np
sys.path.insert(0,'.')
from pymoc_b200 import configs
from pymoc_b200.ensemble import Ensemble
import torch
for name, spec in (('C4', configs.c4_jansen_nadeau(32768)), ('C5', configs.c5_single_global_basin(4096))):
    ens = Ensemble(spec)
    for n in (2400, 4800):
        torch.cuda.synchronize(); t = time.time()
        ens.run(n)
        torch.cuda.synchronize(); dt = time.time() - t
        d = ens.diagnostics(); st = d['status']
        print(name, 'after', ens.it, 'time %.3fs' % dt, 'rate %.3e' % (spec.M * n / dt), 'status bits', {b: int(((st & b) != 0).sum()) for b in (1, 2, 4, 8, 16)}, flush=True)
        bad = (st & 1) != 0
        if bad.any():
            for k, v in spec.sweep.items():
                vals = np.unique(v)
                print('   ', k, ['%.3g:%.2f' % (x, float(bad[v == x].mean())) for x in vals])
