"""Drop-in ``SO_ML`` (reference: src/pymoc/modules/SO_ML.py) on the GPU.

``timestep`` / ``advdiff`` run pmoc_ml_timestep: the surface streamfunction by ``np.interp``
semantics (including numpy's guess-carrying search when ``b_basin`` is not monotone), upwind
advection, surface flux + restoring, and the Crank-Nicolson diffusion as a tridiagonal solve
instead of the reference's dense ``np.linalg.inv`` (SO_ML.py:191-196).  Like the reference,
``advdiff`` *rebinds* ``self.bs`` (SO_ML.py:259,269) rather than mutating it in place.
"""
import numpy as np

from .. import _abi
from ..utils import make_array
from ._dispatch import Call, byref


class SO_ML(object):

  def __init__(self, y=None, Ks=0., h=50., L=4e6, surflux=0., rest_mask=0., b_rest=0., v_pist=1.5 / 86400., bs=0.0,
               Psi_s=None):
    if isinstance(y, np.ndarray):
      self.y = y
    else:
      raise TypeError('y needs to be numpy array providing (regular) grid')
    self.Ks = Ks
    self.h = h
    self.L = L
    self.surflux = make_array(surflux, self.y, 'surflux')
    self.rest_mask = make_array(rest_mask, self.y, 'rest_mask')
    self.b_rest = make_array(b_rest, self.y, 'b_rest')
    self.v_pist = v_pist
    self.Psi_s = Psi_s
    self.bs = make_array(bs, self.y, 'bs')

  def advdiff(self, b_basin, Psi_b, dt):
    y = np.ascontiguousarray(self.y, dtype=np.float64)
    bb = np.ascontiguousarray(b_basin, dtype=np.float64)
    pb = np.ascontiguousarray(Psi_b, dtype=np.float64)
    if not np.any(pb != 0):
      np.nonzero(pb)[0][0]  # the reference's IndexError (SO_ML.py:229)
    if y.size > _abi.MAX_NY_ML or bb.size > _abi.MAX_NZ_WARP:
      raise ValueError('pymoc_b200.modules.SO_ML: the per-method kernel takes ny <= %d surface points and nz <= %d levels '
                       '(got %d, %d); larger models run through pymoc_b200.ensemble.Ensemble'
                       % (_abi.MAX_NY_ML, _abi.MAX_NZ_WARP, y.size, bb.size))
    c = Call()
    m = _abi.Model()
    m.M, m.nz, m.ny = 1, bb.size, y.size
    m.y = c.ptr(y)
    dev_bs = c.dev(np.asarray(self.bs, dtype=np.float64).reshape(1, y.size))
    psi_s = c.out((1, y.size))
    m.ml_bs, m.ml_Psi_s = c.be.ptr(dev_bs), c.be.ptr(psi_s)
    m.ml_Ks, m.ml_h, m.ml_L, m.ml_vpist = (c.vec(float(v)) for v in (self.Ks, self.h, self.L, self.v_pist))
    ones = 0 * y + 1.
    m.ml_surflux, m.ml_rest_mask, m.ml_b_rest = (c.vec(np.asarray(v, dtype=np.float64) * ones)
                                                 for v in (self.surflux, self.rest_mask, self.b_rest))
    status = c.out((1,), np.uint32)
    c.check(c.lib.pmoc_ml_timestep(byref(m), c.vec(bb), c.vec(pb), float(dt), c.be.ptr(status), c.be.stream()))
    st = int(c.get(status).view(np.uint32)[0])
    if st & _abi.ST_ML_INDEX:
      raise IndexError('index 0 is out of bounds for axis 0 with size 0')
    self.Psi_s = c.get(psi_s)[0]
    self.bs = c.get(dev_bs)[0]

  def timestep(self, b_basin=None, Psi_b=None, dt=1.):
    if not isinstance(b_basin, np.ndarray):
      raise TypeError('b_basin needs to be numpy array providing buoyancy levels in basin')
    if not isinstance(Psi_b, np.ndarray):
      raise TypeError('Psi_b needs to be numpy array providing overturning at buoyancy levels given by b_basin')
    self.advdiff(b_basin=b_basin, Psi_b=Psi_b, dt=dt)
