"""Host-side description of an ensemble of coupled PyMOC models.

A :class:`ModelSpec` is what the reference's example scripts build implicitly with
their module instances and hand-written loop: which modules
exist, how they are coupled, and every parameter -- here with a leading *member* axis so
that thousands to millions of independent configurations can be stepped in lock-step.
Per ``north_star`` every user callable is sampled on the grid on the host, once, and
only arrays reach the GPU.

Broadcasting rule: any per-member array may have a leading dimension of 1 instead of
``M``; it is then shared by all members (member stride 0 on the device).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np


def _vec(v, name):
  """Per-member scalar parameter -> contiguous 1-D float64 array (length 1 or M)."""
  a = np.atleast_1d(np.asarray(v, dtype=np.float64))
  if a.ndim != 1:
    raise ValueError('%s must be a scalar or a 1-D per-member array' % name)
  return np.ascontiguousarray(a)


def _prof(v, n, name):
  """Per-member profile -> contiguous [1 or M, n] float64 array."""
  a = np.asarray(v, dtype=np.float64)
  if a.ndim == 0:
    a = np.full((1, n), float(a))
  elif a.ndim == 1:
    a = a[None, :]
  if a.ndim != 2 or a.shape[1] != n:
    raise ValueError('%s must have trailing dimension %d' % (name, n))
  return np.ascontiguousarray(a)


@dataclass
class ColumnSpec:
  """One advective-diffusive column (reference ``Column``, column.py:19-72).

  ``kappa`` carries ``nvar`` alternative diffusivity profiles: the Jansen & Nadeau loop
  re-assigns ``column.kappa`` between two callables every step
  (examples/run_JansenNadeau_2018.py:233-254); variant 0 is the full profile, variant 1
  the bottom-boundary-layer tapered one.  All other loops use ``nvar == 1``.
  """
  kappa: np.ndarray  # [1|M, nvar, nz]
  Area: np.ndarray  # [1|M, nz]
  bs: np.ndarray  # [1|M]
  bbot: np.ndarray  # [1|M]
  N2min: np.ndarray  # [1|M]
  b0: np.ndarray  # [1|M, nz] initial buoyancy
  do_conv: bool = False
  bzbot: Optional[np.ndarray] = None  # [1|M]; when given it replaces bbot (column.py:232-233)
  var0: int = 0  # kappa variant in use at t=0

  @staticmethod
  def build(z, kappa, Area, bs, b0, bbot=0.0, bzbot=None, N2min=1e-7, do_conv=False, var0=0):
    nz = len(z)
    k = np.asarray(kappa, dtype=np.float64)
    if k.ndim == 0:
      k = np.full((1, 1, nz), float(k))
    elif k.ndim == 1:
      k = k[None, None, :]
    elif k.ndim == 2:
      k = k[:, None, :]
    return ColumnSpec(
        kappa=np.ascontiguousarray(k), Area=_prof(Area, nz, 'Area'), bs=_vec(bs, 'bs'),
        bbot=_vec(bbot, 'bbot'), N2min=_vec(N2min, 'N2min'), b0=_prof(b0, nz, 'b'),
        do_conv=bool(do_conv), bzbot=None if bzbot is None else _vec(bzbot, 'bzbot'), var0=int(var0))

  def dAkappa_dz(self, z):
    """``np.gradient(Area*kappa, z)`` per member and variant (column.py:122)."""
    return np.gradient(self.Area[:, None, :] * self.kappa, z, axis=-1)


@dataclass
class ThermwindSpec:
  """Thermal-wind closure (reference ``Psi_Thermwind``, psi_thermwind.py:30-70)."""
  f: np.ndarray  # [1|M]
  b2: Optional[np.ndarray] = None  # [1|M, nz] fixed northern profile when there is no north column

  @staticmethod
  def build(z, f=1.2e-4, b2=None):
    return ThermwindSpec(f=_vec(f, 'f'), b2=None if b2 is None else _prof(b2, len(z), 'b2'))


@dataclass
class ChannelSpec:
  """Southern-Ocean residual circulation (reference ``Psi_SO``, psi_SO.py:17-104).

  ``tau`` is one float per member (every example does this) or an array on ``y`` per
  member.  The four taper heights are shared by the ensemble.

  Shapes: a 0-D / 1-D ``tau`` is read as one float per member (length 1 or M); a wind stress profile on the
  y grid must be 2-D, ``[1, ny]`` (shared) or ``[M, ny]`` -- the reference accepts a bare ``[ny]`` array for
  its single model, here that would be ambiguous with M == ny members and is rejected by ``ModelSpec``.
  """
  y: np.ndarray  # [ny]
  bs: np.ndarray  # [1|M, ny]
  tau: np.ndarray  # [1|M] or [1|M, ny]
  f: np.ndarray
  rho: np.ndarray
  L: np.ndarray
  KGM: np.ndarray
  smax: np.ndarray
  c: Optional[np.ndarray] = None  # [1|M]; None -> explicit GM branch (psi_SO.py:325-327)
  bvp_with_Ek: bool = False
  Hsill: Optional[float] = None
  HEk: Optional[float] = None
  Htapertop: Optional[float] = None
  Htaperbot: Optional[float] = None

  @staticmethod
  def build(y, bs, tau, f=1.2e-4, rho=1030., L=1e7, KGM=1e3, c=None, bvp_with_Ek=False, Hsill=None,
            HEk=None, Htapertop=None, Htaperbot=None, smax=0.01):
    y = np.ascontiguousarray(np.asarray(y, dtype=np.float64))
    t = np.asarray(tau, dtype=np.float64)
    # 0-D / 1-D: one float per member;  2-D: one profile on y per member
    tau_arr = _vec(t, 'tau') if t.ndim < 2 else _prof(t, y.size, 'tau')
    return ChannelSpec(
        y=y, bs=_prof(bs, y.size, 'bs'), tau=tau_arr, f=_vec(f, 'f'), rho=_vec(rho, 'rho'),
        L=_vec(L, 'L'), KGM=_vec(KGM, 'KGM'), smax=_vec(smax, 'smax'),
        c=None if c is None else _vec(c, 'c'), bvp_with_Ek=bool(bvp_with_Ek), Hsill=Hsill, HEk=HEk,
        Htapertop=Htapertop, Htaperbot=Htaperbot)


@dataclass
class MixedLayerSpec:
  """Southern-Ocean mixed layer (reference ``SO_ML``, SO_ML.py:17-71)."""
  y: np.ndarray  # [ny] uniform
  Ks: np.ndarray
  h: np.ndarray
  L: np.ndarray
  v_pist: np.ndarray
  surflux: np.ndarray  # [1|M, ny]
  rest_mask: np.ndarray
  b_rest: np.ndarray
  bs0: np.ndarray  # [1|M, ny] initial surface buoyancy

  @staticmethod
  def build(y, bs, Ks=0., h=50., L=4e6, surflux=0., rest_mask=0., b_rest=0., v_pist=1.5 / 86400.):
    y = np.ascontiguousarray(np.asarray(y, dtype=np.float64))
    n = y.size
    return MixedLayerSpec(
        y=y, Ks=_vec(Ks, 'Ks'), h=_vec(h, 'h'), L=_vec(L, 'L'), v_pist=_vec(v_pist, 'v_pist'),
        surflux=_prof(surflux, n, 'surflux'), rest_mask=_prof(rest_mask, n, 'rest_mask'),
        b_rest=_prof(b_rest, n, 'b_rest'), bs0=_prof(bs, n, 'bs'))


@dataclass
class ModelSpec:
  """An ensemble of ``M`` coupled models sharing grids, time step and topology.

  order 'post': step the columns, then re-diagnose the streamfunctions on iterations
                with ``ii % K == 0`` (examples/example_timestepping.py:73-80,
                example_twocol.py:85-96, example_twocol_plusSO.py:99-115).
  order 'jn'  : re-diagnose at the top of iterations with ``ii % K == 0``, apply the
                bottom-boundary switches, step both columns with convection, then the
                mixed layer (examples/run_JansenNadeau_2018.py:201-261,
                run_single_global_basin.py:172-229).
  iso         : columns are forced with the isopycnally remapped streamfunction
                (``Psibz``) instead of the z-space one.
  pac         : the two-basin topology of examples/twobasin_NadeauJansen.py:63-122 -- a third
                column coupled to ``basin`` through a second thermal-wind closure (``zoc_f``,
                the zonal overturning in the channel) and to the channel through a second
                ``Psi_SO`` that differs from ``so`` in its zonal length ``so_pac_L`` only.
  """
  M: int
  z: np.ndarray
  dt: float
  K: int
  basin: ColumnSpec
  north: Optional[ColumnSpec] = None
  tw: Optional[ThermwindSpec] = None
  so: Optional[ChannelSpec] = None
  ml: Optional[MixedLayerSpec] = None
  pac: Optional[ColumnSpec] = None
  zoc_f: Optional[np.ndarray] = None  # [1|M]
  so_pac_L: Optional[np.ndarray] = None  # [1|M]
  order: str = 'post'
  iso: bool = False
  nb: int = 500
  name: str = ''
  sweep: dict = field(default_factory=dict)  # free-form: the lattice axes, for reports

  def __post_init__(self):
    self.z = np.ascontiguousarray(np.asarray(self.z, dtype=np.float64))
    if self.order not in ('post', 'jn'):
      raise ValueError("order must be 'post' or 'jn'")
    if self.order == 'jn' and not (self.north and self.tw and self.so and self.ml and self.iso):
      raise ValueError("order 'jn' needs basin + north + tw (iso) + so + ml")
    if self.iso and not self.tw:
      raise ValueError('iso needs a thermal-wind closure')
    if self.north is not None and not (self.tw and self.iso):
      raise ValueError('a north column is coupled through the isopycnal thermal-wind closure')
    if self.tw is not None and self.north is None and self.tw.b2 is None:
      raise ValueError('tw.b2 is needed when there is no north column')
    if self.pac is not None:
      if not (self.north and self.tw and self.iso and self.so) or self.ml is not None or self.order != 'post':
        raise ValueError("the two-basin topology needs basin + north + tw (iso) + so, order 'post', no mixed layer")
      if self.zoc_f is None or self.so_pac_L is None:
        raise ValueError('pac needs zoc_f and so_pac_L')
      self.zoc_f, self.so_pac_L = _vec(self.zoc_f, 'zoc_f'), _vec(self.so_pac_L, 'so_pac_L')

    self._check_member_axes()

  def _check_member_axes(self):
    """Every per-member array must have a leading dimension of 1 (shared by all members) or M.  A short array
    would otherwise be uploaded as it is and read out of bounds on the device for members beyond its length
    (e.g. a wind stress given on the ny-point y grid as a 1-D array: that must be passed as [1, ny])."""
    nz, M = self.z.size, int(self.M)

    def chk(owner, name, a, trailing):
      if a is None:
        return
      a = np.asarray(a)
      if a.ndim != 1 + len(trailing) or tuple(a.shape[1:]) != tuple(trailing) or a.shape[0] not in (1, M):
        raise ValueError('%s.%s: shape %r, expected (1|%d%s)' % (owner, name, a.shape, M, ''.join(', %d' % t for t in trailing)))

    for tag, c in (('basin', self.basin), ('north', self.north), ('pac', self.pac)):
      if c is None:
        continue
      if c.kappa.ndim != 3 or c.kappa.shape[2] != nz or c.kappa.shape[0] not in (1, M) or c.kappa.shape[1] not in (1, 2):
        raise ValueError('%s.kappa: shape %r, expected (1|%d, 1|2 variants, %d); a [nvar, nz] array must be given as '
                         '[1, nvar, nz]' % (tag, c.kappa.shape, M, nz))
      if not 0 <= c.var0 < c.kappa.shape[1]:
        raise ValueError('%s.var0 = %d outside the %d kappa variant(s)' % (tag, c.var0, c.kappa.shape[1]))
      for name, tr in (('Area', (nz,)), ('b0', (nz,)), ('bs', ()), ('bbot', ()), ('N2min', ()), ('bzbot', ())):
        chk(tag, name, getattr(c, name), tr)
    if self.tw is not None:
      chk('tw', 'f', self.tw.f, ())
      chk('tw', 'b2', self.tw.b2, (nz,))
    if self.so is not None:
      so, ny = self.so, self.so.y.size
      chk('so', 'bs', so.bs, (ny,))
      chk('so', 'tau', so.tau, (ny,) if so.tau.ndim == 2 else ())
      for name in ('f', 'rho', 'L', 'KGM', 'smax', 'c'):
        chk('so', name, getattr(so, name), ())
    if self.ml is not None:
      ml, ny = self.ml, self.ml.y.size
      if self.so is not None and ny != self.so.y.size:
        raise ValueError('ml.y and so.y must be the same grid')
      for name in ('Ks', 'h', 'L', 'v_pist'):
        chk('ml', name, getattr(ml, name), ())
      for name in ('surflux', 'rest_mask', 'b_rest', 'bs0'):
        chk('ml', name, getattr(ml, name), (ny,))
    chk('model', 'zoc_f', self.zoc_f, ())
    chk('model', 'so_pac_L', self.so_pac_L, ())

  @property
  def nz(self):
    return self.z.size

  @property
  def ny(self):
    return 0 if self.so is None else self.so.y.size

  # ------------------------------------------------------------------ one member
  def member_case(self, m):
    """Member ``m`` as the plain dict the CPU oracle consumes (oracle/pymoc_oracle.py)."""
    pick = lambda a: a[m if a.shape[0] > 1 else 0]

    def col(c):
      return dict(kappa=pick(c.kappa).copy(), Area=pick(c.Area).copy(), bs=float(pick(c.bs)),
                  bbot=float(pick(c.bbot)), bzbot=None if c.bzbot is None else float(pick(c.bzbot)),
                  N2min=float(pick(c.N2min)), b0=pick(c.b0).copy(), do_conv=c.do_conv, var0=c.var0)

    case = dict(z=self.z.copy(), dt=float(self.dt), K=int(self.K), nb=int(self.nb), order=self.order,
                iso=bool(self.iso), basin=col(self.basin),
                north=None if self.north is None else col(self.north), tw=None, so=None, ml=None)
    if self.pac is not None:
      case['pac'] = col(self.pac)
      case['zoc_f'] = float(pick(self.zoc_f))
      case['so_pac_L'] = float(pick(self.so_pac_L))
    if self.tw is not None:
      case['tw'] = dict(f=float(pick(self.tw.f)), b2=None if self.tw.b2 is None else pick(self.tw.b2).copy())
    if self.so is not None:
      s = self.so
      tau = pick(s.tau)
      case['so'] = dict(y=s.y.copy(), bs=pick(s.bs).copy(), tau=float(tau) if np.ndim(tau) == 0 else tau.copy(),
                        f=float(pick(s.f)), rho=float(pick(s.rho)), L=float(pick(s.L)), KGM=float(pick(s.KGM)),
                        smax=float(pick(s.smax)), c=None if s.c is None else float(pick(s.c)),
                        bvp_with_Ek=s.bvp_with_Ek, Hsill=s.Hsill, HEk=s.HEk, Htapertop=s.Htapertop,
                        Htaperbot=s.Htaperbot)
    if self.ml is not None:
      l = self.ml
      case['ml'] = dict(y=l.y.copy(), Ks=float(pick(l.Ks)), h=float(pick(l.h)), L=float(pick(l.L)),
                        v_pist=float(pick(l.v_pist)), surflux=pick(l.surflux).copy(),
                        rest_mask=pick(l.rest_mask).copy(), b_rest=pick(l.b_rest).copy(), bs=pick(l.bs0).copy())
    return case
