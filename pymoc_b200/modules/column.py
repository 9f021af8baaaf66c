"""Drop-in ``Column`` (reference: src/pymoc/modules/column.py) running on the GPU.

Same constructor keywords and defaults (column.py:19-29), same attributes, same in-place
mutation of ``self.b`` (column.py:231-232,249,268,271,312), same exceptions.  Every
``timestep`` / ``vertadvdiff`` / ``convect`` / ``horadv`` call samples the callables on the
grid on the host and runs ``pmoc_column_timestep`` (include/pymoc_b200.h).  The equilibrium
solver (``solve_equi``/``ode``/``bc``) is outside the time-stepping hot path: a host-side utility
that, like the reference, hands its one boundary value problem to ``scipy.integrate.solve_bvp``.
"""
import numpy as np

from .. import _abi
from ..utils import check_numpy_version, make_array, make_func
from ._dispatch import Call, byref


class Column(object):

  def __init__(self, z=None, kappa=None, bs=0.025, bbot=0.0, bzbot=None, b=0.0, Area=None, N2min=1e-7):
    if isinstance(z, np.ndarray) and len(z) > 0:
      self.z = z
    else:
      raise TypeError('z needs to be numpy array providing grid levels')
    self.kappa = make_func(kappa, self.z, 'kappa')
    self.Area = make_func(Area, self.z, 'Area')
    self.bs = bs
    self.bbot = bbot
    self.bzbot = bzbot
    self.N2min = N2min
    self.b = make_array(b, self.z, 'b')
    self.bz = np.gradient(self.b, z) if check_numpy_version() else 0. * z

  # --- host-side, state independent (column.py:74-122) ---------------------------------------
  def Akappa(self, z):
    return self.Area(z) * self.kappa(z)

  def dAkappa_dz(self, z):
    if not check_numpy_version():
      raise ImportError('You need NumPy version 1.13.0 or later. Please upgrade your NumPy libary.')
    return np.gradient(self.Akappa(z), z)

  # --- equilibrium profile (column.py:124-208).  Not time stepping: one two-point
  # boundary value problem per call, which the reference hands to scipy.integrate.solve_bvp.  It is a host-side
  # set-up utility here as well (same solver, same ODE, same boundary residuals), for API completeness
  # (examples/example_iteration.py:63); nothing on the GPU path calls it.
  def bc(self, ya, yb):
    """Boundary residuals: bottom buoyancy (or its gradient when ``bzbot`` is given) and surface buoyancy."""
    if self.bzbot is None:
      return np.array([ya[0] - self.bbot, yb[0] - self.bs])
    return np.array([ya[1] - self.bzbot, yb[0] - self.bs])

  def ode(self, z, y):
    """b' = y1;  y1' = (wA - d(A kappa)/dz) / (A kappa) * y1  (steady advection-diffusion, column.py:155-185)."""
    return np.vstack((y[1], (self.wA(z) - self.dAkappa_dz(z)) / self.Akappa(z) * y[1]))

  def solve_equi(self, wA):
    """Equilibrium buoyancy profile for the area-integrated velocity ``wA`` (column.py:187-208).  Like the
    reference it REBINDS ``self.b`` / ``self.bz`` to new arrays."""
    from scipy import integrate
    self.wA = make_func(wA, self.z, 'w')
    guess = np.zeros((2, np.size(self.z)))
    guess[0, :] = self.b
    guess[1, :] = self.bz
    res = integrate.solve_bvp(self.ode, self.bc, self.z, guess)
    self.b = res.sol(self.z)[0, :]
    self.bz = res.sol(self.z)[1, :]

  # --- the GPU call ---------------------------------------------------------------------------
  def _run(self, stages, wA=None, dt=1., do_conv=False, vdx_in=None, b_in=None):
    z = np.ascontiguousarray(self.z, dtype=np.float64)
    nz = z.size
    if nz > _abi.MAX_NZ_WARP:
      raise ValueError('pymoc_b200.modules.Column: the per-method kernels hold one column per warp, nz <= %d (got %d); '
                       'taller columns run through pymoc_b200.ensemble.Ensemble' % (_abi.MAX_NZ_WARP, nz))
    c = Call()
    col = _abi.Column()
    dev_b = c.dev(np.asarray(self.b, dtype=np.float64).reshape(1, nz))
    col.b = c.be.ptr(dev_b)
    ones = 0 * z + 1.
    col.kappa = c.vec(self.kappa(z) * ones)
    col.dAk = c.vec(self.dAkappa_dz(z))
    col.Area = c.vec(self.Area(z) * ones)
    col.bs = c.vec(self.bs)
    col.N2min = c.vec(self.N2min)
    col.bzbot = c.vec(self.bzbot)
    col.bbot = c.ptr(np.array([self.bbot], dtype=np.float64))
    col.nvar, col.do_conv = 1, int(bool(do_conv))
    vwA = c.vec(wA) if wA is not None else _abi.Vec(None, 0)
    vv = c.vec(vdx_in) if vdx_in is not None else _abi.Vec(None, 0)
    vb = c.vec(b_in) if b_in is not None else _abi.Vec(None, 0)
    c.check(c.lib.pmoc_column_timestep(1, nz, c.ptr(z), byref(col), vwA, vv, vb, float(dt), stages, c.be.stream()))
    self.b[...] = c.get(dev_b)[0]  # in place: callers alias this array (make_array.py:30-31)

  def vertadvdiff(self, wA, dt, do_conv=False):
    wA = make_array(wA, self.z, 'wA')
    self._run(_abi.STAGE_VERTADVDIFF, wA=wA, dt=dt, do_conv=do_conv)

  def convect(self):
    self._run(_abi.STAGE_CONVECT)

  def horadv(self, vdx_in, b_in, dt):
    vdx_in = make_array(vdx_in, self.z, 'vdx_in')
    b_in = make_array(b_in, self.z, 'b_in')
    self._run(_abi.STAGE_HORADV, dt=dt, vdx_in=vdx_in, b_in=b_in)

  def timestep(self, wA=0., dt=1., do_conv=False, vdx_in=None, b_in=None):
    stages = _abi.STAGE_VERTADVDIFF | (_abi.STAGE_CONVECT if do_conv else 0)
    missing = vdx_in is not None and b_in is None
    if vdx_in is not None and not missing:
      stages |= _abi.STAGE_HORADV
      vdx_in = make_array(vdx_in, self.z, 'vdx_in')
      b_in = make_array(b_in, self.z, 'b_in')
    else:
      vdx_in = b_in = None
    self._run(stages, wA=make_array(wA, self.z, 'wA'), dt=dt, do_conv=do_conv, vdx_in=vdx_in, b_in=b_in)
    if missing:  # like the reference, AFTER convect + vertadvdiff have changed self.b (column.py:336-348)
      raise TypeError('b_in is needed if vdx_in is provided')
