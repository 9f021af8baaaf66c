"""CPU oracle for the PyMOC time-stepping hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain NumPy/SciPy and one member at a time, the algorithm of the
reference classes on the path named by BASELINE.json.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it; the product package ``pymoc_b200`` never does.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function here
against fixtures under ``tests/golden/`` that were produced by importing and running the
unmodified reference from ``/root/reference`` (``tests/golden/make_golden.py``).

The third-party numerics the reference reaches (``scipy.integrate.solve_bvp``,
``scipy.optimize.brentq``, ``np.interp``, ``np.gradient``, ``np.linalg.inv``) are present
in this image (numpy 2.3.5 / scipy 1.18.1) and are called here exactly where the
reference calls them, so the default modes are bit-for-bit the reference's arithmetic.
Three alternative modes state the *closed forms the CUDA kernels use* so that the gap
between them and the reference's iterative solvers can be measured on the CPU:

  thermwind='quad'   exact double cumulative quadrature of Psi'' = (b2-b1)/f
                     (the reference's own comment, psi_thermwind.py:131-132)
  ys='brentq_restated'  Brent's method written out (same iterates as scipy's C code),
                     'analytic' = first linear crossing north of argmin(bs)
  cn='thomas'        tridiagonal solve instead of the dense np.linalg.inv

All citations are ``file:line`` under ``/root/reference/src/pymoc/modules`` unless
prefixed otherwise.
"""
from __future__ import annotations

import numpy as np
from scipy import integrate, optimize

SV = 1e6  # m^3/s per Sverdrup


# --------------------------------------------------------------------------- Column
class ColumnState:
  """Plain container for one advective-diffusive column (column.py:19-72)."""

  def __init__(self, z, kappa, Area, b, bs, bbot=0.0, bzbot=None, N2min=1e-7):
    self.z = np.asarray(z, dtype=np.float64)
    self.b = np.array(b, dtype=np.float64)  # own copy: the oracle never aliases
    self.bs, self.bbot, self.bzbot, self.N2min = bs, bbot, bzbot, N2min
    self.set_kappa(kappa, Area)

  def set_kappa(self, kappa, Area=None):
    """(Re)sample kappa / Area on the grid; d(A kappa)/dz as column.py:122."""
    if Area is not None:
      self.Area = np.asarray(Area, dtype=np.float64) + 0 * self.z
    self.kappa = np.asarray(kappa, dtype=np.float64) + 0 * self.z
    self.dAk = np.gradient(self.Area * self.kappa, self.z)


def column_convect(col: ColumnState):
  """Downward convective adjustment, column.py:264-271 (strict ``>``)."""
  unstable = col.b > col.bs
  if unstable.any():
    stable = ~unstable
    z_anchor = col.z[stable].max() if stable.any() else col.z[0]
    col.b[unstable] = col.bs + col.N2min * (col.z[unstable] - z_anchor)
  else:
    col.b[-1] = col.bs


def column_vertadvdiff(col: ColumnState, wA, dt, do_conv=False):
  """Explicit upwind advection + diffusion step, column.py:226-249."""
  z, b = col.z, col.b
  wA = np.asarray(wA, dtype=np.float64) + 0 * z
  h = z[1:] - z[:-1]
  if not do_conv:
    b[-1] = col.bs
  b[0] = col.bbot if col.bzbot is None else b[1] - col.bzbot * h[0]
  slope = (b[1:] - b[:-1]) / h
  above, below = slope[1:], slope[:-1]
  curvature = (above - below) / (0.5 * (h[1:] + h[:-1]))
  w = (wA - col.dAk)[1:-1]
  upwind = np.where(w < 0, above, below)
  tendency = -w * upwind / col.Area[1:-1] + col.kappa[1:-1] * curvature
  b[1:-1] = b[1:-1] + dt * tendency


def column_horadv(col: ColumnState, vdx_in, b_in, dt):
  """Upwind lateral inflow, column.py:306-313."""
  vdx_in = np.asarray(vdx_in, dtype=np.float64) + 0 * col.z
  b_in = np.asarray(b_in, dtype=np.float64) + 0 * col.z
  inflow = vdx_in > 0.0
  delta = b_in - col.b
  col.b[inflow] = col.b[inflow] + dt * vdx_in[inflow] * delta[inflow] / col.Area[inflow]


def column_timestep(col, wA=0., dt=1., do_conv=False, vdx_in=None, b_in=None):
  """Operation order of column.py:336-348."""
  if do_conv:
    column_convect(col)
  column_vertadvdiff(col, wA, dt, do_conv)
  if vdx_in is not None:
    if b_in is None:
      raise TypeError('b_in is needed if vdx_in is provided')
    column_horadv(col, vdx_in, b_in, dt)


# ------------------------------------------------------------------ Psi_Thermwind
def thermwind_solve(z, b1, b2, f, method='bvp'):
  """Psi'' = (b2-b1)/f with Psi(z0)=Psi(zN)=0, in Sv (psi_thermwind.py:123-135).

  'bvp'  : scipy's collocation solver, cold-started from zeros like the reference.
  'quad' : the exact solution for piecewise-linear b's (a C1 piecewise cubic):
           I1 = cumulative trapezoid of g, I2 = cumulative exact integral of I1,
           Psi = I2 - I2[-1] (z-z0)/(zN-z0).
  """
  z = np.asarray(z, dtype=np.float64)
  b1 = np.asarray(b1, dtype=np.float64) + 0 * z
  b2 = np.asarray(b2, dtype=np.float64) + 0 * z
  if method == 'bvp':
    rhs = lambda x, y: np.vstack((y[1], 1. / f * (np.interp(x, z, b2) - np.interp(x, z, b1))))
    ends = lambda ya, yb: np.array([ya[0], yb[0]])
    res = integrate.solve_bvp(rhs, ends, z, np.zeros((2, z.size)))
    return res.sol(z)[0, :] / SV
  g = 1. / f * (b2 - b1)
  h = z[1:] - z[:-1]
  first = np.concatenate(([0.], np.cumsum(0.5 * (g[1:] + g[:-1]) * h)))
  cell = first[:-1] * h + g[:-1] * (h * h) / 2. + (g[1:] - g[:-1]) * (h * h) / 6.
  second = np.concatenate(([0.], np.cumsum(cell)))
  return (second - second[-1] * ((z - z[0]) / (z[-1] - z[0]))) / SV


def thermwind_psib(Psi, b1, b2, nb=500):
  """Upwind isopycnal remap, psi_thermwind.py:170-185.  Returns (psib, bgrid).

  The (nb, nz-1) matrix form below performs, row by row, the very operations of the
  reference's ``for i in range(nb)`` loop (element-wise IEEE ops, then a pairwise
  ``np.sum`` along the contiguous axis), inf/NaN behaviour of flat cells included.
  """
  b1 = np.asarray(b1, dtype=np.float64)
  b2 = np.asarray(b2, dtype=np.float64) + 0 * b1
  bgrid = np.linspace(min(b1.min(), b2.min()), max(b1.max(), b2.max()), nb)
  transport = -(Psi[1:] - Psi[:-1])
  from_b2 = transport < 0
  lower = np.where(from_b2, b2[:-1], b1[:-1])
  upper = np.where(from_b2, b2[1:], b1[1:])
  with np.errstate(divide='ignore', invalid='ignore'):
    psib = np.empty(nb)
    for i in range(nb):
      frac = np.clip((upper - bgrid[i]) / (upper - lower), 0., 1.)
      psib[i] = np.sum(frac * transport)
  return psib, bgrid


def thermwind_psibz(Psi, b1, b2, nb=500):
  """psi_thermwind.py:203-208: psib interpolated back onto each column's b(z)."""
  b1 = np.asarray(b1, dtype=np.float64)
  b2 = np.asarray(b2, dtype=np.float64) + 0 * b1
  psib, bgrid = thermwind_psib(Psi, b1, b2, nb)
  return np.interp(b1, bgrid, psib), np.interp(b2, bgrid, psib), psib, bgrid


# ------------------------------------------------------------------------- Psi_SO
class ChannelParams:
  """Constants of Psi_SO (psi_SO.py:17-104).  ``tau`` is a float or an array on y."""

  def __init__(self, z, y, tau, f=1.2e-4, rho=1030, L=1e7, KGM=1e3, c=None,
               bvp_with_Ek=False, Hsill=None, HEk=None, Htapertop=None, Htaperbot=None,
               smax=0.01):
    self.z = np.asarray(z, dtype=np.float64)
    self.y = np.asarray(y, dtype=np.float64)
    self.tau, self.f, self.rho, self.L, self.KGM = tau, f, rho, L, KGM
    self.c, self.bvp_with_Ek, self.smax = c, bvp_with_Ek, smax
    self.Hsill, self.HEk, self.Htapertop, self.Htaperbot = Hsill, HEk, Htapertop, Htaperbot


def brentq_restated(f, xa, xb, xtol=2e-12, rtol=8.881784197001252e-16, maxiter=100):
  """Plain restatement of SciPy's Brent root finder (scipy 1.18.1,
  ``scipy/optimize/Zeros/brentq.c``; published algorithm: Brent 1973, ch. 4) with
  ``optimize.brentq``'s default tolerances.  The CUDA kernel carries the same
  statement sequence, so that for a multi-root (non-monotone) bs(y) it lands on the
  very root the reference finds.  Checked bit-for-bit against
  ``scipy.optimize.brentq`` in tests/test_oracle_golden.py.
  """
  xpre, xcur = xa, xb
  xblk = fblk = spre = scur = 0.
  fpre, fcur = f(xpre), f(xcur)
  if fpre == 0:
    return xpre
  if fcur == 0:
    return xcur
  if np.signbit(fpre) == np.signbit(fcur):
    raise ValueError('f(a) and f(b) must have different signs')
  for _ in range(maxiter):
    if fpre != 0 and fcur != 0 and np.signbit(fpre) != np.signbit(fcur):
      xblk, fblk = xpre, fpre
      spre = scur = xcur - xpre
    if abs(fblk) < abs(fcur):
      xpre, xcur, xblk = xcur, xblk, xcur
      fpre, fcur, fblk = fcur, fblk, fcur
    delta = (xtol + rtol * abs(xcur)) / 2
    sbis = (xblk - xcur) / 2
    if fcur == 0 or abs(sbis) < delta:
      return xcur
    if abs(spre) > delta and abs(fcur) < abs(fpre):
      if xpre == xblk:
        stry = -fcur * (xcur - xpre) / (fcur - fpre)
      else:
        dpre = (fpre - fcur) / (xpre - xcur)
        dblk = (fblk - fcur) / (xblk - xcur)
        stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre))
      if 2 * abs(stry) < min(abs(spre), 3 * abs(sbis) - delta):
        spre, scur = scur, stry
      else:
        spre = scur = sbis
    else:
      spre = scur = sbis
    xpre, fpre = xcur, fcur
    if abs(scur) > delta:
      xcur += scur
    else:
      xcur += delta if sbis > 0 else -delta
    fcur = f(xcur)
  return xcur


def so_outcrop(bval, y, bs, method='brentq'):
  """Outcrop latitude ys(b), psi_SO.py:125-140."""
  if bval < bs.min():
    return y[0] - 1e3
  if bval > bs[-1]:
    return y[-1]
  south = int(np.argmin(bs))
  if method == 'brentq':
    return optimize.brentq(lambda yy: np.interp(yy, y, bs) - bval, y[south], y[-1])
  if method == 'brentq_restated':
    return brentq_restated(lambda yy: float(np.interp(yy, y, bs)) - bval, y[south], y[-1])
  # analytic: the bracket ends are returned when they are exact roots (brentq does the
  # same), otherwise the first sign change north of argmin(bs) is inverted linearly.
  if bs[south] == bval:
    return y[south]
  if bs[-1] == bval:
    return y[-1]
  for j in range(south, y.size - 1):
    lo, hi = bs[j] - bval, bs[j + 1] - bval
    if hi == 0.0:
      return y[j + 1]
    if lo * hi < 0.0:
      return y[j] + (bval - bs[j]) * (y[j + 1] - y[j]) / (bs[j + 1] - bs[j])
  return y[-1]


def _bottom_taper(H, z):
  return 1. if H is None else 1. - np.maximum(z[0] + H - z, 0.)**2. / H**2.


def _top_taper(H, z, scalar=True):
  if H is not None:
    return 1 - np.maximum(z + H, 0)**2. / H**2.
  if scalar:
    return 1.
  ones = np.ones(z.size)
  ones[-1] = 0.
  return ones


def so_tau_average(p: ChannelParams, y0):
  """100-point mean of tau between the outcrop and the northern edge, psi_SO.py:239."""
  pts = np.linspace(y0, p.y[-1], 100)
  if isinstance(p.tau, np.ndarray):
    return np.mean(np.interp(pts, p.y, p.tau))
  return np.mean(p.tau + 0 * pts)


def so_ekman(p: ChannelParams, b, bs, ys='brentq'):
  """Psi_Ek in m^3/s, psi_SO.py:236-243."""
  tau_ave = np.array([so_tau_average(p, so_outcrop(bi, p.y, bs, ys)) for bi in b])
  return tau_ave / p.f / p.rho * p.L * _bottom_taper(p.Hsill, p.z) * _top_taper(p.HEk, p.z, False)


def so_n2(z, b):
  """Centred / one-sided stratification, psi_SO.py:154-160."""
  h = z[1:] - z[:-1]
  n2 = np.zeros(z.size)
  n2[1:-1] = (b[2:] - b[:-2]) / (h[1:] + h[:-1])
  n2[0] = (b[1] - b[0]) / h[0]
  n2[-1] = (b[-1] - b[-2]) / h[-1]
  return n2


def so_gm(p: ChannelParams, b, bs, Psi_Ek_sv, ys='brentq'):
  """Psi_GM in m^3/s, psi_SO.py:302-331 (both the explicit and the F2010 BVP branch)."""
  z, y = p.z, p.y
  width = np.array([max(y[-1] - so_outcrop(bi, y, bs, ys), 0.1) for bi in b])
  tapers_bot, tapers_top = _bottom_taper(p.Htaperbot, z), _top_taper(p.Htapertop, z)
  if p.c is not None:
    target = p.KGM * z / width * p.L * tapers_top * tapers_bot
    n2 = so_n2(z, b)
    rhs = lambda x, s: np.vstack((s[1], np.interp(x, z, n2) / p.c**2. * (s[0] - np.interp(x, z, target))))
    if p.bvp_with_Ek:
      ends = lambda sa, sb: np.array([sa[0] + Psi_Ek_sv[0] * 1e6, sb[0] + Psi_Ek_sv[-1] * 1e6])
    else:
      ends = lambda sa, sb: np.array([sa[0], sb[0]])
    gm = integrate.solve_bvp(rhs, ends, z, np.zeros((2, z.size))).sol(z)[0, :]
  else:
    gm = p.KGM * np.maximum(z / width, -p.smax) * p.L * tapers_top * tapers_bot
  blocked = width > y[-1] - y[0]
  gm[blocked] = np.maximum(gm[blocked], -Psi_Ek_sv[blocked] * 1e6)
  return gm


def so_solve(p: ChannelParams, b, bs, ys='brentq'):
  """psi_SO.py:349-354.  Returns (Psi, Psi_Ek, Psi_GM) in Sv."""
  b = np.asarray(b, dtype=np.float64) + 0 * p.z
  bs = np.asarray(bs, dtype=np.float64) + 0 * p.y
  ek = so_ekman(p, b, bs, ys) / 1e6
  gm = so_gm(p, b, bs, ek, ys) / 1e6
  psi = ek + gm
  psi[0] = 0.
  return psi, ek, gm


# -------------------------------------------------------------------------- SO_ML
class MixedLayerState:
  """Southern-Ocean mixed layer, SO_ML.py:17-71."""

  def __init__(self, y, Ks=0., h=50., L=4e6, surflux=0., rest_mask=0., b_rest=0.,
               v_pist=1.5 / 86400., bs=0.0):
    self.y = np.asarray(y, dtype=np.float64)
    grid = lambda v: np.array(np.asarray(v, dtype=np.float64) + 0 * self.y)
    self.Ks, self.h, self.L, self.v_pist = Ks, h, L, v_pist
    self.surflux, self.rest_mask, self.b_rest = grid(surflux), grid(rest_mask), grid(b_rest)
    self.bs = grid(bs)
    self.Psi_s = None


def _ml_south_bc(ml, b_basin, Psi_b):
  """SO_ML.py:93-98."""
  if ml.Psi_s[1] > 0:
    ml.bs[0] = b_basin[np.argwhere(Psi_b > 0)[0][0]]
  else:
    ml.bs[0] = ml.bs[1]


def _cn_matrix(n, s):
  """SO_ML.py:155-165."""
  mat = (np.diag(-s / 2. * np.ones(n - 1), -1) + np.diag((1 + s) * np.ones(n), 0) +
         np.diag(-s / 2. * np.ones(n - 1), 1))
  mat[0, 0], mat[0, 1], mat[-1, -2], mat[-1, -1] = 1, 0, 0, 1
  return mat


def ml_diffuse(bs, s, cn='inv'):
  """Crank-Nicolson diffusion, SO_ML.py:191-196."""
  n = bs.size
  if cn == 'inv':
    return np.dot(np.dot(np.linalg.inv(_cn_matrix(n, s)), _cn_matrix(n, -s)), bs)
  rhs = bs.copy()
  rhs[1:-1] = s / 2. * bs[:-2] + (1 - s) * bs[1:-1] + s / 2. * bs[2:]
  # Thomas sweep on tridiag(-s/2, 1+s, -s/2) with identity first/last rows
  cp, dp = np.zeros(n), np.zeros(n)
  dp[0] = rhs[0]
  for i in range(1, n - 1):
    m = 1. / ((1 + s) + (s / 2.) * cp[i - 1])
    cp[i] = -(s / 2.) * m
    dp[i] = (rhs[i] + (s / 2.) * dp[i - 1]) * m
  out = np.empty(n)
  out[-1] = rhs[-1]
  for i in range(n - 2, 0, -1):
    out[i] = dp[i] - cp[i] * out[i + 1]
  out[0] = rhs[0]
  return out


def ml_timestep(ml: MixedLayerState, b_basin, Psi_b, dt, cn='inv'):
  """SO_ML.py:228-274."""
  held = Psi_b.copy()
  first = np.nonzero(held)[0][0]
  held[:first] = held[first]
  ml.Psi_s = np.interp(ml.bs, b_basin, held)
  ml.Psi_s[:np.argmin(ml.bs)] = 0.
  ml.Psi_s[0] = 0.
  _ml_south_bc(ml, b_basin, Psi_b)
  flux = ml.surflux / ml.h + ml.rest_mask * ml.v_pist / ml.h * (ml.b_rest - ml.bs)
  dy = ml.y[1] - ml.y[0]
  adv = 0. * ml.y
  inner = ml.Psi_s[1:-1]
  neg, pos = inner < 0., inner > 0.
  adv[1:-1][neg] = -inner[neg] * 1e6 * (ml.bs[2:][neg] - ml.bs[1:-1][neg]) / ml.h / ml.L / dy
  adv[1:-1][pos] = -inner[pos] * 1e6 * (ml.bs[1:-1][pos] - ml.bs[:-2][pos]) / ml.h / ml.L / dy
  ml.bs = ml.bs + dt * (flux + adv)
  if ml.Psi_s[1] <= 0:
    ml.bs[0] = ml.bs[1]
  ml.bs = ml_diffuse(ml.bs, ml.Ks * dt / dy**2, cn)
  _ml_south_bc(ml, b_basin, Psi_b)


# ---------------------------------------------------------------- coupling loops
class Modes:
  """Which closed forms replace the reference's iterative solvers (see module doc)."""

  def __init__(self, thermwind='bvp', ys='brentq', cn='inv'):
    self.thermwind, self.ys, self.cn = thermwind, ys, cn


REFERENCE = Modes()
KERNEL = Modes('quad', 'brentq_restated', 'thomas')


def run_coupled(case, nsteps, modes=REFERENCE, it0=0, carry=None):
  """Advance one member of a coupled model by ``nsteps``.

  ``case`` is the dict layout written by ``tests/golden/make_golden.py`` (one member):
    z, dt, K, nb, order ('post' | 'jn'), iso (bool)
    basin / north : dict(kappa[nvar,nz], Area[nz], bs, bbot, bzbot|None, N2min, b0[nz], do_conv)
                    ('north' may be absent; then tw['b2'] is the fixed northern profile)
    tw            : dict(f, b2) or None
    so            : dict(ChannelParams kwargs + 'bs' array) or None
    ml            : dict(MixedLayerState kwargs) or None

  order 'post' : step the columns, then refresh the streamfunctions when ii % K == 0
                 (examples/example_timestepping.py:73-80, example_twocol.py:85-96,
                  example_twocol_plusSO.py:99-115);
  order 'jn'   : refresh at the top of iteration ii % K == 0, apply the bottom-boundary
                 switches, step the columns, then the mixed layer
                 (examples/run_JansenNadeau_2018.py:201-261,
                  examples/run_single_global_basin.py:172-229).

  Returns a dict with the final state and the last diagnosed streamfunctions.
  """
  z = np.asarray(case['z'], dtype=np.float64)
  dt, K, nb = case['dt'], int(case['K']), int(case.get('nb', 500))

  def make_col(d):
    col = ColumnState(z, d['kappa'][int(d.get('var0', 0))], d['Area'], d['b0'], d['bs'], d['bbot'],
                      d.get('bzbot'), d['N2min'])
    col.variants = [(np.asarray(k, dtype=np.float64), np.gradient(col.Area * np.asarray(k), z))
                    for k in d['kappa']]
    col.do_conv = bool(d['do_conv'])
    return col

  def use_variant(col, v):
    col.kappa, col.dAk = col.variants[v]

  basin = make_col(case['basin'])
  north = make_col(case['north']) if case.get('north') is not None else None
  pac = make_col(case['pac']) if case.get('pac') is not None else None  # examples/twobasin_NadeauJansen.py:88
  tw, so, ml = case.get('tw'), case.get('so'), case.get('ml')
  chan = ChannelParams(z, so['y'], so['tau'], **{k: so[k] for k in so if k not in ('y', 'tau', 'bs')}) if so else None
  bs_chan = np.array(so['bs'], dtype=np.float64) if so else None
  chan_pac = None
  if pac is not None:  # SO_Pac: same surface buoyancy, wind and eddy parameters, its own zonal length (:80-81)
    chan_pac = ChannelParams(z, so['y'], so['tau'],
                             **{**{k: so[k] for k in so if k not in ('y', 'tau', 'bs')}, 'L': case['so_pac_L']})
  layer = MixedLayerState(**ml) if ml else None
  if carry is not None:  # resume (pickup) support: state only, Psi's are re-diagnosed
    basin.b[:] = carry['b_basin']
    if north is not None:
      north.b[:] = carry['b_north']
    if layer is not None:
      layer.bs[:] = carry['bs_ml']
    if pac is not None:
      pac.b[:] = carry['b_pac']

  out = {}

  def refresh():
    if tw:
      b2 = north.b if north is not None else np.asarray(tw['b2'], dtype=np.float64) + 0 * z
      out['Psi_tw'] = thermwind_solve(z, basin.b, b2, tw['f'], modes.thermwind)
      if case.get('iso', False):
        out['Psi_iso_b'], out['Psi_iso_n'], out['psib'], out['bgrid'] = thermwind_psibz(
            out['Psi_tw'], basin.b, b2, nb)
    if pac is not None:  # ZOC between the two basins (:117-119)
      out['Psi_zoc'] = thermwind_solve(z, basin.b, pac.b, case['zoc_f'], modes.thermwind)
      out['Psi_zon_a'], out['Psi_zon_p'], out['psib2'], out['bgrid2'] = thermwind_psibz(out['Psi_zoc'], basin.b, pac.b, nb)
    if so:
      surf = layer.bs if layer is not None else bs_chan
      out['Psi_so'], out['Psi_Ek'], out['Psi_GM'] = so_solve(chan, basin.b, surf, modes.ys)
      if pac is not None:
        out['Psi_so2'], out['Psi_Ek2'], out['Psi_GM2'] = so_solve(chan_pac, pac.b, surf, modes.ys)

  def velocities():
    north_leg = (out['Psi_iso_b'] if case.get('iso', False) else out['Psi_tw']) if tw else 0. * z
    south_leg = out['Psi_so'] if so else 0. * z
    wAb = (north_leg - south_leg) * 1e6
    wAn = -out['Psi_iso_n'] * 1e6 if (tw and north is not None) else None
    return wAb, wAn

  if case['order'] == 'post':
    if it0 == 0 or carry is None:
      refresh()
    else:
      out.update(carry['psi'])
    for ii in range(it0, it0 + nsteps):
      wAb, wAn = velocities()
      if pac is not None:  # examples/twobasin_NadeauJansen.py:104-109
        wAb = (out['Psi_iso_b'] + out['Psi_zon_a'] - out['Psi_so']) * 1e6
        wAp = (-out['Psi_zon_p'] - out['Psi_so2']) * 1e6
      column_timestep(basin, wAb, dt, basin.do_conv)
      if north is not None:
        column_timestep(north, wAn, dt, north.do_conv)
      if pac is not None:
        column_timestep(pac, wAp, dt, pac.do_conv)
      if ii % K == 0:
        refresh()
  else:
    for ii in range(it0, it0 + nsteps):
      if ii % K == 0:
        refresh()
      wAb, wAn = velocities()
      psi_so, res_b, res_n = out['Psi_so'], out['Psi_iso_b'], out['Psi_iso_n']
      # run_JansenNadeau_2018.py:233-254 -- bottom boundary / BBL-kappa switches
      if psi_so[1] < 0:
        basin.bbot = layer.bs[0]
        use_variant(basin, 1)
      if res_b[1] > 0 and north.b[0] < basin.b[1] and north.b[0] < layer.bs[0]:
        basin.bbot = north.b[0]
        use_variant(basin, 1)
      elif psi_so[1] >= 0:
        basin.bbot = basin.b[1]
        use_variant(basin, 0)
      if res_n[1] < 0 and basin.b[0] < north.b[1]:
        north.bbot = basin.b[0]
        use_variant(north, 1)
      else:
        north.bbot = north.b[1]
        use_variant(north, 0)
      column_timestep(basin, wAb, dt, True)
      column_timestep(north, wAn, dt, True)
      ml_timestep(layer, basin.b, psi_so, dt, modes.cn)

  out['b_basin'] = basin.b.copy()
  if north is not None:
    out['b_north'] = north.b.copy()
  if pac is not None:
    out['b_pac'] = pac.b.copy()
  if layer is not None:
    out['bs_ml'] = layer.bs.copy()
    out['Psi_s'] = None if layer.Psi_s is None else layer.Psi_s.copy()
  return out
