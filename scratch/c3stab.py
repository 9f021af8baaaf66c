import sys, numpy as np
sys.path.insert(0,'.')
from pymoc_b200 import configs
from pymoc_b200.ensemble import Ensemble
spec = configs.c3_twocol_so(32768, axes=(16,16,16,8))
ens = Ensemble(spec)
for n in (2400, 7200, 14400):
    ens.run(n)
    d = ens.diagnostics()
    st = d['status']
    bad = (st & 1) != 0
    print('after', ens.it, 'NaN members', int(bad.sum()), 'maxPsi_tw', float(np.nanmax(np.abs(d['Psi_tw'][~bad]))), flush=True)
    if bad.any():
        for k, v in spec.sweep.items():
            vals = np.unique(v)
            frac = [float(bad[v == x].mean()) for x in vals]
            print(' ', k, ['%.3g:%.2f' % (x, f) for x, f in zip(vals, frac)])
