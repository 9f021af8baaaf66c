"""C3_bvp long runs: GPU kernel vs CPU warp emulator (same sources) vs oracle, same 32 lattice members"""
import sys, warnings
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
warnings.filterwarnings('ignore')
from parity_common import sample_cases, oracle_many
from helpers import spec_from_cases, relmax
from pymoc_b200.ensemble import Ensemble
from emu.emu_backend import EmuBackend
if __name__ == '__main__':
  ms, cases = sample_cases('C3_bvp', 262144, 32, 20261019)
  spec = spec_from_cases(cases)
  g = Ensemble(spec); e = Ensemble(spec, backend=EmuBackend())
  keys = ('b_basin', 'b_north', 'Psi_so', 'Psi_iso_b')
  for n in (2400, 4800, 7200):
    g.run(n - g.it); e.run(n - e.it)
    G = {**g.state(), **g.diagnostics()}; E = {**e.state(), **e.diagnostics()}
    want = oracle_many(cases, n)
    ge = np.array([max(relmax(G[k][i], E[k][i]) for k in keys) for i in range(len(ms))])
    go = np.array([max(relmax(G[k][i], want[i][k]) for k in keys) for i in range(len(ms))])
    eo = np.array([max(relmax(E[k][i], want[i][k]) for k in keys) for i in range(len(ms))])
    print(n, 'gpu vs emu worst %.2e (member %d) | gpu vs oracle worst %.2e (member %d) | emu vs oracle worst %.2e (member %d)' % (
        ge.max(), ms[ge.argmax()], go.max(), ms[go.argmax()], eo.max(), ms[eo.argmax()]), flush=True)
    print('   gpu vs oracle sorted top', np.sort(go)[-4:], ' emu vs oracle sorted top', np.sort(eo)[-4:], flush=True)
