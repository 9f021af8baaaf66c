#!/usr/bin/env python
"""The same two-column + Southern-Ocean model as a lock-step ensemble: a tau x kappa x bs_north x A_basin
lattice (pymoc_b200.configs.c3_twocol_so) stepped by ONE fused kernel launch per call, every member's state
resident on chip.  This replaces the hand-written ``for ii in range(total_iters)`` loop of
examples/example_twocol_plusSO.py:99-115 for thousands to millions of members.

    python examples/ensemble_sweep.py --members 4096 --iters 2400
"""
import argparse
import time

import numpy as np

from pymoc_b200 import configs
from pymoc_b200.ensemble import Ensemble


def main(members, iters):
  spec = configs.c3_twocol_so(members)      # c=None: the explicit-GM twin; c=0.1 for the script's F2010 smoother
  ens = Ensemble(spec)
  t0 = time.perf_counter()
  ens.run(iters)                            # pmoc_model_diagnose + pmoc_model_run
  dt = time.perf_counter() - t0
  psi = ens.diagnostics()['Psi_tw']         # [M, nz] Sv
  amoc = psi.max(axis=1)
  i = int(amoc.argmax())
  print('%d members x %d steps in %.3f s (%.3g member-steps/s)' % (members, iters, dt, members * iters / dt))
  print('strongest AMOC %.2f Sv at tau=%.3f kappa=%.2e bs_north=%.4f A_basin=%.2e' %
        (amoc[i], spec.sweep['tau'][i], spec.sweep['kappa'][i], spec.sweep['bs_north'][i], spec.sweep['A_basin'][i]))
  return amoc


if __name__ == '__main__':
  ap = argparse.ArgumentParser()
  ap.add_argument('--members', type=int, default=4096)
  ap.add_argument('--iters', type=int, default=2400)
  a = ap.parse_args()
  main(a.members, a.iters)
