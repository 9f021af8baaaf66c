import sys, time, numpy as np
sys.path.insert(0,'.')
from pymoc_b200 import configs
from pymoc_b200.ensemble import Ensemble
import torch
def rate(spec, nt, reps=3):
    ens = Ensemble(spec)
    ens.run(nt)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ens.run(nt, sync=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = ens.diagnostics()['status']
    return spec.M * nt / ms * 1e3, ms, {b: int(((st & b) != 0).sum()) for b in (1, 2, 4, 8, 16, 32, 64)}
for name, mk, nt in (('C2', lambda: configs.c2_column_so(65536), 7200), ('C3', lambda: configs.c3_twocol_so(32768), 2400), ('C4', lambda: configs.c4_jansen_nadeau(32768), 2400), ('C5', lambda: configs.c5_single_global_basin(32768), 2400), ('C1', lambda: configs.c1_timestepping(16384), 3000)):
    spec = mk()
    r, ms, st = rate(spec, nt)
    print(name, 'K=%d' % spec.K, 'rate %.4e' % r, 'ms %.2f' % ms, st, flush=True)
    if name in ('C2', 'C3', 'C4'):
        spec = mk(); spec.K = 1000000
        r2, ms2, st = rate(spec, nt)
        print(name, 'K=inf', 'rate %.4e' % r2, 'ms %.2f' % ms2, ' => refresh share %.1f%%' % (100 * (1 - ms2 / ms)), flush=True)
