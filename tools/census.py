"""Status-bit census of the bench lattices at bench size and length (one fused launch each): how many members
carry each PMOC_ST_* bit, i.e. for how many the reference's own answer is decided by rounding noise.
usage: python tools/census.py [WORKLOAD ...] > gpurun_out/census.json"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from pymoc_b200 import _abi  # noqa: E402
from pymoc_b200.ensemble import Ensemble  # noqa: E402

out = {}
for wl in (sys.argv[1:] or ['C1', 'C2', 'C3', 'C3_bvp', 'twobasin', 'C4', 'C5', 'C5_4096']):
  build, M, nt = bench.WORKLOADS[wl]
  if wl == 'C5_4096':
    M = 2048
  spec = build(M)
  ens = Ensemble(spec)
  t = time.time()
  ens.run(nt)
  torch.cuda.synchronize()
  dt = time.time() - t
  c = _abi.status_census(ens.diagnostics()['status'])
  out[wl] = dict(members=M, steps=nt, seconds=round(dt, 3), census=c)
  print(wl, out[wl], file=sys.stderr, flush=True)
  del ens
print(json.dumps(out))
