"""Pickup and diagnostics files in the reference's wire format.

``examples/run_JansenNadeau_2018.py`` (and ``run_single_global_basin.py``) read and write two
kinds of ``np.savez`` archives with *positional* array names:

* pickup (``--pickup`` / ``--pickup_save_file``, :61-64, :135-138, :266-267)::

      arr_0 = basin.b[nz]   arr_1 = north.b[nz]   arr_2 = channel.bs[ny]

  The streamfunctions are re-diagnosed from the state at iteration 0, so this is a complete
  checkpoint provided the iteration counter restarts at a multiple of ``MOC_up_iters``.

* diagnostics (``--diagfile``, :191-198, :218-226, :268-272), one column per ``Diag_iters``::

      arr_0 AMOC.Psi[nz, nd]   arr_1 AMOC.Psib(nb)[nb, nd]   arr_2 basin.b[nz, nd]   arr_3 north.b[nz, nd]
      arr_4 channel.bs[ny, nd] arr_5 z   arr_6 AMOC.bgrid[nb, nd]   arr_7 y   arr_8 PsiSO.Psi[nz, nd]
      arr_9 tau   arr_10 kapGM

  which is what ``examples/Plot_overturning.py:21-31`` consumes.

Here both exist per member (the reference's exact layout, so its plotting scripts stay drop-in)
and for the whole ensemble (the same positional names with a leading member axis).
"""
from __future__ import annotations

import numpy as np


def save_pickup(ens, path, member=None):
  """``np.savez(path, basin.b, north.b, channel.bs)``; ``member=None`` keeps the member axis."""
  st = ens.state()
  arrs = [st['b_basin'], st['b_north'], st['bs_ml']]
  if member is not None:
    arrs = [a[member] for a in arrs]
  np.savez(path, *arrs)


def load_pickup(ens, path, member=None):
  """Set the ensemble's state from a pickup archive.  A single-member archive (1-D arrays) is
  broadcast to every member (or written to ``member`` only); a batched one must match ``ens.M``.
  The iteration counter restarts at 0, as the scripts' does."""
  f = np.load(path)
  new = {'b_basin': f['arr_0'], 'b_north': f['arr_1'], 'bs_ml': f['arr_2']}
  cur = ens.state()
  for k, v in new.items():
    v = np.asarray(v, dtype=np.float64)
    if v.ndim == 1:
      if member is None:
        cur[k][:] = v[None, :]
      else:
        cur[k][member] = v
    else:
      if v.shape != cur[k].shape:
        raise ValueError('%s: pickup holds %r, the ensemble %r' % (k, v.shape, cur[k].shape))
      cur[k] = v
  ens.set_state(**cur)
  ens.it = 0


class DiagRecorder:
  """Runs the loop of run_JansenNadeau_2018.py:201-261 and keeps what the script saves on
  iterations with ``ii % Diag_iters == 0`` (:218-226): the freshly diagnosed AMOC.Psi, AMOC.Psib,
  bgrid and PsiSO.Psi together with the state at the top of that iteration."""

  KEYS = ('Psi_tw', 'psib', 'b_basin', 'b_north', 'bs_ml', 'bgrid', 'Psi_so')

  def __init__(self, ens, diag_iters):
    if ens.spec.order != 'jn':
      raise ValueError("the diagnostics file is written by the 'jn' scripts")
    if diag_iters % ens.spec.K != 0:
      raise ValueError('Diag_iters must be a multiple of MOC_up_iters (the script samples inside the refresh)')
    self.ens, self.diag_iters = ens, int(diag_iters)
    self.cols = {k: [] for k in self.KEYS}

  def run(self, total_iters):
    ens = self.ens
    if ens.it % self.diag_iters != 0:
      raise ValueError('start the recorder on a diagnostics iteration')
    for _ in range(int(total_iters // self.diag_iters)):
      ens.diagnose()  # what the refresh at the top of this iteration computes; the state is untouched
      got = {**ens.state(), **ens.diagnostics()}
      for k in self.KEYS:
        self.cols[k].append(got[k])
      ens.run(self.diag_iters)
    return self

  def arrays(self, member=None):
    """The 11 positional arrays; time is the last axis as in the script."""
    spec = self.ens.spec
    stack = lambda k: np.stack(self.cols[k], axis=-1)  # [M, n, nd]
    pick = (lambda a: a) if member is None else (lambda a: a[member])
    par = lambda a: a if member is None else a[member if a.shape[0] > 1 else 0]
    so = spec.so
    lo = self.ens.lo
    sl = lambda a: a if a.shape[0] == 1 else a[lo:lo + self.ens.M]
    return [pick(stack('Psi_tw')), pick(stack('psib')), pick(stack('b_basin')), pick(stack('b_north')),
            pick(stack('bs_ml')), spec.z.copy(), pick(stack('bgrid')), so.y.copy(), pick(stack('Psi_so')),
            par(sl(so.tau)), par(sl(so.KGM))]

  def save(self, path, member=None):
    np.savez(path, *self.arrays(member))
