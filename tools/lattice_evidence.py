"""One-off parity evidence at the BENCH's own length: seeded-random members of the bench-size lattices, stepped for
the number of steps one bench launch runs, every one against the live oracle (tests/parity_common.lattice_sample:
unflagged members must match, flagged ones are counted).  Appends one JSON line per workload to the report.
usage: python tools/lattice_evidence.py <report.jsonl> [workload:members:steps ...]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
os.environ['PMOC_LATTICE_REPORT'] = os.path.abspath(sys.argv[1])

import bench  # noqa: E402
from parity_common import lattice_sample  # noqa: E402
from pymoc_b200.backend import CudaBackend  # noqa: E402

if __name__ == '__main__':
  cuda = CudaBackend()
  jobs = sys.argv[2:] or ['C4:512:2400', 'C5:512:2400', 'C3:512:2400', 'twobasin:256:2400', 'C2:256:7200', 'C1:128:3000',
                          'C3_bvp:64:2400']
  for job in jobs:
    wl, n, steps, *opt = job.split(':')
    K = next((int(o[1:]) for o in opt if o.startswith('K')), None)       # e.g. C5_4096:8:45:K20
    tol_opt = next((float(o[1:]) for o in opt if o.startswith('T')), None)  # e.g. C3_bvp:32:7200:T1e-3
    M = {'C3': 262144, 'C3_bvp': 262144, 'C4': 1048576}.get(wl, bench.WORKLOADS[wl][1])  # BASELINE sizes
    t = time.time()
    rep = lattice_sample(cuda, wl, M, int(n), int(steps), seed=20261019,
                         tol=tol_opt if tol_opt is not None else (1e-5 if wl == 'C3_bvp' else 1e-10), K=K)
    print(wl, 'lattice', M, 'sampled', rep['sampled'], 'steps', steps, 'worst unflagged %.2e' % rep['worst_unflagged'],
          'flagged', rep['census']['parity_undefined'], 'matching', rep['flagged_matching'], 'missing', rep['flagged_missing'],
          'lost by both', rep['lost_by_both'], '%.0f s' % (time.time() - t), flush=True)
