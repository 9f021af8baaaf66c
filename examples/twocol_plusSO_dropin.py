#!/usr/bin/env python
"""The two-column + Southern-Ocean model of the reference's examples/example_twocol_plusSO.py, run through
the drop-in classes: the only change against the reference script is the import line
(``from pymoc_b200.modules import ...`` instead of ``from pymoc.modules import ...``); every method call below
is one kernel launch through the C ABI.  (Plotting is left out; the loop statements are the script's,
example_twocol_plusSO.py:99-115.)

    python examples/twocol_plusSO_dropin.py --iters 240

This is the one-member, launch-per-call way; examples/ensemble_sweep.py runs the same model as a lock-step
ensemble in one fused kernel.
"""
import argparse

import numpy as np

from pymoc_b200.modules import Column, Psi_SO, Psi_Thermwind


def main(iters, c=0.1, nodal=False):
  """nodal=True hands the initial profiles to Psi_Thermwind as arrays on z instead of callables (the script's
  callables are also sampled at the cell mid-points by the BVP solve; the batched engine is nodal)."""
  bs, bs_north, bmin = 0.03, 0.004, 0.0
  A_basin, A_north = 6e13, 6e13 / 50.
  kappa = 2e-5
  y = np.asarray(np.linspace(0, 2.e6, 40))
  bs_SO = (bs - bmin) * (y / y[-1])**2 + bmin
  dt = 86400 * 30
  MOC_up_iters = int(np.floor(2. * 360 * 86400 / dt))
  z = np.asarray(np.linspace(-4000, 0, 80))
  b_basin = lambda zz: bs * np.exp(zz / 300.)
  b_north = lambda zz: bs_north * np.exp(zz / 300.)

  AMOC = Psi_Thermwind(z=z, b1=b_basin(z) if nodal else b_basin, b2=b_north(z) if nodal else b_north, f=1e-4)
  AMOC.solve()
  Psi_iso_b, Psi_iso_n = AMOC.Psibz()
  SO = Psi_SO(z=z, y=y, b=b_basin(z), bs=bs_SO, tau=0.13, f=1e-4, L=5e6, KGM=1000., c=c, bvp_with_Ek=c is not None)
  SO.solve()
  basin = Column(z=z, kappa=kappa, Area=A_basin, b=b_basin, bs=bs, bbot=bmin)
  north = Column(z=z, kappa=kappa, Area=A_north, b=b_north, bs=bs_north, bbot=bmin)

  for ii in range(iters):
    wAb = (Psi_iso_b - SO.Psi) * 1e6
    wAN = -Psi_iso_n * 1e6
    basin.timestep(wA=wAb, dt=dt)
    north.timestep(wA=wAN, dt=dt, do_conv=True)
    if ii % MOC_up_iters == 0:
      AMOC.update(b1=basin.b, b2=north.b)
      AMOC.solve()
      Psi_iso_b, Psi_iso_n = AMOC.Psibz()
      SO.update(b=basin.b)
      SO.solve()
  return dict(b_basin=basin.b, b_north=north.b, Psi=AMOC.Psi, Psi_SO=SO.Psi)


if __name__ == '__main__':
  ap = argparse.ArgumentParser()
  ap.add_argument('--iters', type=int, default=240)
  out = main(ap.parse_args().iters)
  print('AMOC max %.3f Sv at z = %.0f m; SO min %.3f Sv' % (out['Psi'].max(), np.linspace(-4000, 0, 80)[out['Psi'].argmax()],
                                                          out['Psi_SO'].min()))
