#!/usr/bin/env python
"""examples/run_JansenNadeau_2018.py as a lock-step ensemble whose arrays stay in HOST memory: the script's loop with
its diagnostics every ``Diag_iters`` iterations (run_JansenNadeau_2018.py:201-226, 266-272), through the persistent
host-buffer handle of the C ABI (pmoc_host_open / pmoc_host_step / pmoc_host_close).  Grids and parameters go to
the GPU once; every call advances ``Diag_iters`` iterations and brings back the state and the streamfunctions, which
are then read from the (pinned) host arrays exactly where the script fills its ``*_save`` arrays.

    python examples/jansen_nadeau_host_loop.py --members 4096 --iters 1200 --pickup-save pickup.npz
"""
import argparse
import time

import numpy as np

from pymoc_b200 import _abi, configs, pickup
from pymoc_b200.ensemble import HostEnsemble


def main(members, iters, diag_iters=120, backend=None, pickup_save=None):
  spec = configs.c4_jansen_nadeau(members)   # tau x kapfac x db x B x KGM lattice around the script's own values
  ens = HostEnsemble(spec, backend=backend)
  S, P, D = ens.IO_STATE, ens.IO_PSI, ens.IO_DIAG
  amoc = []                                   # max of the isopycnal overturning at every diagnostic time, per member
  t0 = time.perf_counter()
  for _ in range(0, iters, diag_iters):
    ens.run(diag_iters, pull=S | P | D)       # nothing goes up: the device copy is current
    amoc.append(ens.diagnostics()['psib'].max(axis=1))
  dt = time.perf_counter() - t0
  census = _abi.status_census(ens.diagnostics()['status'])
  print('%d members x %d steps in %.3f s (%.3g member-steps/s), %d diagnostic times' % (members, iters, dt, members * iters / dt, len(amoc)))
  print('members whose reference answer hangs on rounding noise: %d of %d %s' % (
      census['parity_undefined'], census['members'], {k: v for k, v in census.items() if v and k not in ('members', 'clean')}))
  if pickup_save:
    pickup.save_pickup(ens, pickup_save)      # arr_0, arr_1, arr_2 = b_basin, b_north, bs_SO with a leading member axis
  ens.close()
  return np.array(amoc)


if __name__ == '__main__':
  ap = argparse.ArgumentParser()
  ap.add_argument('--members', type=int, default=4096)
  ap.add_argument('--iters', type=int, default=1200)
  ap.add_argument('--diag-iters', type=int, default=120)
  ap.add_argument('--pickup-save', default=None)
  a = ap.parse_args()
  main(a.members, a.iters, a.diag_iters, pickup_save=a.pickup_save)
