"""Checks of the drop-in classes, shared by the emulator (CPU) and CUDA (GPU) test files.

They mirror the reference's own unit tests (tests/modules/test_column.py etc.) and add the
bit-level fixtures of tests/golden/units.npz produced by the reference itself.
"""
import numpy as np
import pytest

from helpers import golden, relmax
from pymoc_b200.modules import SO_ML, Column, Psi_SO, Psi_Thermwind

TOL = 1e-10
TOL_BVP = 1e-5  # Psi_SO with c != None


# ------------------------------------------------------------------------------- Column
def column_init_errors():
  z = np.asarray(np.linspace(-4000, 0, 80))
  with pytest.raises(TypeError) as e:
    Column(z=z, kappa=2e-5, Area=None)
  assert str(e.value) == "('Area', 'needs to be either function, numpy array, or float')"
  with pytest.raises(TypeError) as e:
    Column(z=z, kappa=None, Area=6e13)
  assert str(e.value) == "('kappa', 'needs to be either function, numpy array, or float')"
  for bad in (None, 50, np.array([])):
    with pytest.raises(TypeError) as e:
      Column(z=bad, kappa=2e-5, Area=6e13)
    assert str(e.value) == 'z needs to be numpy array providing grid levels'
  col = Column(z=z, kappa=2e-5, Area=6e13)
  assert (col.bs, col.bbot, col.bzbot, col.N2min) == (0.025, 0.0, None, 1e-7)
  assert np.array_equal(col.b, 0. * z)
  arr = np.linspace(0.03, -0.001, 80)
  assert Column(z=z, kappa=2e-5, Area=6e13, b=arr).b is arr  # aliasing, make_array.py:30-31


def column_reference_tests():
  # tests/modules/test_column.py:272-282 (exact)
  N2min = 1.5e-7
  z = np.asarray([-4000.0, -1000.0, -100.0, 0.0])
  b = np.asarray([-0.03, -0.02, 0.01, 0.01])
  col = Column(z=z, b=b.copy(), bs=0.0, N2min=N2min, kappa=2e-5, Area=6e13)
  b[2:] = 0.0 + N2min * (z[2:] - z[1])
  col.convect()
  assert all(col.b == b)
  # test_column.py:247-270 (3 decimals)
  dt, Area = 60 * 86400, 6e13
  z = np.asarray(np.linspace(-4000, 0, 80))
  b = np.linspace(0.03, -0.002, 80)
  wA = Area * np.sin(z)
  db_dt1 = (0.0004 / 50.0 / Area) * (-wA)
  col = Column(z=z, Area=Area, kappa=2e-5, b=b.copy(), bbot=-0.002, bs=0.03)
  col.vertadvdiff(wA, dt, do_conv=False)
  assert all(np.around(col.b[2:-2], decimals=3) == np.around(b[2:-2] - dt * db_dt1[2:-2], decimals=3))
  # test_column.py:284-296 (4 decimals)
  z = np.asarray([-4000.0, -1000.0, -100.0, 0.0])
  b = np.asarray([-0.03, 0.01, -0.0025, -0.002])
  col = Column(z=z, b=b.copy(), kappa=2e-5, Area=6e13)
  b[0] = -0.03 + dt * 2e6 / 6e13
  col.horadv(np.asarray([2e8, 2.5e8, 0.0, 0.0]), np.asarray([-0.02, 0.01, -0.001, 0.001]), dt)
  assert all(np.around(col.b, decimals=4) == np.around(b, decimals=4))
  # test_column.py:298-336: order convect -> vertadvdiff -> horadv, exact composition
  z = np.asarray(np.linspace(-4000, 0, 80))
  b = np.linspace(-np.sqrt(0.04), 0.0, 80)**2.
  vdx_in = np.asarray([2e4 for n in z])
  b_in = np.asarray([-0.02 for n in z])
  wA = np.sin(z) / Area
  dt = 30 * 86400
  mk = lambda: Column(z=z, b=b.copy(), bs=-0.0, bbot=-0.04, kappa=2e-5, Area=Area)
  c1, c2 = mk(), mk()
  c1.timestep(wA=wA, dt=dt)
  c2.vertadvdiff(wA=wA, dt=dt)
  assert all(c1.b == c2.b)
  c2.horadv(vdx_in=vdx_in, b_in=b_in, dt=dt)
  c2.convect()
  assert any(c1.b != c2.b)
  c1, c2 = mk(), mk()
  c1.timestep(wA=wA, dt=dt, b_in=b_in, vdx_in=vdx_in)
  c2.vertadvdiff(wA=wA, dt=dt)
  c2.horadv(vdx_in=vdx_in, b_in=b_in, dt=dt)
  assert all(c1.b == c2.b)
  c1, c2 = mk(), mk()
  c1.timestep(wA=wA, dt=dt, b_in=b_in, vdx_in=vdx_in, do_conv=True)
  c2.convect()
  c2.vertadvdiff(wA=wA, dt=dt)
  c2.horadv(vdx_in=vdx_in, b_in=b_in, dt=dt)
  assert all(c1.b == c2.b)
  c3, c4 = mk(), mk()
  with pytest.raises(TypeError) as e:
    c3.timestep(wA=wA, dt=dt, vdx_in=vdx_in)
  assert str(e.value) == 'b_in is needed if vdx_in is provided'
  c4.vertadvdiff(wA=wA, dt=dt)  # the reference raises AFTER the column has been stepped (column.py:336-348)
  assert all(c3.b == c4.b)


def column_golden_units():
  for i, d in golden('units')['column'].items():
    inp, out = d['inp'], d['out']
    mk = lambda: Column(z=inp['z'], kappa=inp['kappa'], Area=inp['Area'], b=inp['b'].copy(), bs=inp['bs'],
                        bbot=inp['bbot'], bzbot=inp['bzbot'], N2min=inp['N2min'])
    for tag, kw in (('plain', {}), ('conv', dict(do_conv=True)),
                    ('conv_hor', dict(do_conv=True, vdx_in=inp['vdx_in'], b_in=inp['b_in'])),
                    ('hor', dict(vdx_in=inp['vdx_in'], b_in=inp['b_in']))):
      col = mk()
      for s in range(3):
        col.timestep(wA=inp['wA'], dt=inp['dt'], **kw)
        assert relmax(col.b, out[tag][s]) < TOL, (i, tag, s, relmax(col.b, out[tag][s]))
    col = mk()
    col.convect()
    assert np.array_equal(col.b, out['convect_only']), i  # un-fused arithmetic: bit exact
    assert np.array_equal(col.dAkappa_dz(inp['z']), out['dAkappa_dz'])


# ------------------------------------------------------------------------ Psi_Thermwind
def thermwind_checks():
  with pytest.raises(TypeError) as e:
    Psi_Thermwind(z=1.0, b1=0.1)
  assert str(e.value) == 'z needs to be numpy array providing grid levels'
  for i, d in golden('units')['thermwind'].items():
    inp, out = d['inp'], d['out']
    if inp['z'].size > 256:
      continue
    tw = Psi_Thermwind(z=inp['z'], b1=inp['b1'].copy(), b2=inp['b2'].copy(), f=inp['f'])
    tw.solve()
    assert relmax(tw.Psi, out['Psi']) < TOL, (i, relmax(tw.Psi, out['Psi']))
    if 'psib' in out:
      tw.Psi = out['Psi'].copy()  # isolate the remap
      with np.errstate(all='ignore'):
        psib = tw.Psib(500)
        assert relmax(tw.bgrid, out['bgrid']) == 0.0, i
        assert relmax(psib, out['psib']) < TOL, (i, relmax(psib, out['psib']))
        iso = tw.Psibz(500)
        assert relmax(iso[0], out['iso_b']) < TOL and relmax(iso[1], out['iso_n']) < TOL, i
        assert relmax(tw.Psib(37), out['psib37']) < TOL, i
  # callable profiles (first Psi of example_timestepping.py): the kernel gets mid-point samples, i.e. what
  # solve_bvp's collocation sees on the un-refined mesh.  For this curved profile solve_bvp inserts a
  # node (70 -> 71, rms residual 3e-4 after) and is therefore adaptive; stated tolerance 1e-6 (measured 2.5e-8).
  z = np.asarray(np.linspace(-3500, 0, 70))
  f_b = lambda zz: 0.03 * np.exp(zz / 300.) - 0.0004
  tw = Psi_Thermwind(z=z, b1=f_b)
  tw.solve()
  from scipy import integrate
  ref = integrate.solve_bvp(lambda x, y: np.vstack((y[1], 1. / 1.2e-4 * (0. + 0 * x - f_b(x)))),
                            lambda ya, yb: np.array([ya[0], yb[0]]), z, np.zeros((2, 70))).sol(z)[0, :] / 1e6
  assert relmax(tw.Psi, ref) < 1e-6, relmax(tw.Psi, ref)


# ------------------------------------------------------------------------------- Psi_SO
def so_checks():
  z = np.asarray(np.linspace(-4000, 0, 81))
  y = np.asarray(np.linspace(0, 2.0e6, 51))
  with pytest.raises(TypeError) as e:
    Psi_SO(z=-2000)
  assert str(e.value) == 'z needs to be numpy array providing grid levels'
  with pytest.raises(TypeError) as e:
    Psi_SO(z=z, y=1e6)
  assert str(e.value) == 'y needs to be numpy array providing horizontal grid (or boundaries) of ACC'
  so = Psi_SO(z=z, y=y, b=np.linspace(0.03, -0.001, 81), bs=np.linspace(0.05, 0.10, 51), tau=0.12)
  # tests/modules/test_psi_SO.py:120-133
  assert np.round(so.ys(0.02), decimals=3) == np.round(y[0] - 1e3, decimals=3)
  assert np.round(so.ys(0.2), decimals=3) == np.round(y[-1], decimals=3)
  for i in range(len(y)):
    assert np.round(so.ys(so.bs(y[i])), decimals=3) == np.round(y[i], decimals=3)
  # test_psi_SO.py:144-155 constant wind stress
  ekman = [(so.L * 0.12) / (so.f * so.rho)] * len(z)
  ekman[-1] = 0
  assert all(np.around(ekman, decimals=3) == np.around(so.calc_Ekman(), decimals=3))
  # test_psi_SO.py:371-412
  so.solve()
  assert so.Psi[0] == 0.0
  assert np.array_equal(so.Psi[1:], (so.Psi_Ek + so.Psi_GM)[1:])
  for i, d in golden('units')['so'].items():
    inp, out = d['inp'], d['out']
    kw = {k: inp[k] for k in ('f', 'rho', 'L', 'KGM', 'c', 'bvp_with_Ek', 'Hsill', 'HEk', 'Htapertop', 'Htaperbot',
                              'smax')}
    so = Psi_SO(z=inp['z'], y=inp['y'], b=inp['b'].copy(), bs=inp['bs'].copy(), tau=inp['tau'], **kw)
    so.solve()
    # F2010 smoother: the reference's adaptive solve_bvp (tol=1e-3) vs the converged solution: stated 1e-5
    tol = TOL if inp['c'] is None else TOL_BVP
    for got, key in ((so.Psi, 'Psi'), (so.Psi_Ek, 'Psi_Ek'), (so.Psi_GM, 'Psi_GM')):
      assert relmax(got, out[key]) < tol, (i, key, relmax(got, out[key]))
    if not isinstance(inp['tau'], np.ndarray):  # (tau(y) is averaged from ys(b), which brentq knows to ~1e-9 m only)
      assert np.array_equal(so.Psi_Ek, out['Psi_Ek']), i  # same operations in the same order: bit exact
    ys = np.array([so.ys(v) for v in inp['b'][::8]])
    assert np.abs(ys - out['ys'][::8]).max() < 1e-6, i  # metres; brentq's own tolerance is ~4e-9 m


# -------------------------------------------------------------------------------- SO_ML
def ml_checks():
  with pytest.raises(TypeError) as e:
    SO_ML(y=100)
  assert str(e.value) == 'y needs to be numpy array providing (regular) grid'
  # tests/modules/test_SO_ML.py:92-127
  dt = 60 * 86400
  conf = {'y': np.asarray(np.linspace(0, 2.0e6, 51)), 'Ks': 100, 'h': 50, 'L': 4e6, 'surflux': 5.9e3, 'rest_mask': 0.0,
          'b_rest': 0.0, 'v_pist': 2.0 / 86400.0, 'bs': 0.02}
  b_basin = np.linspace(0.03, -0.002, 80)
  Psi_b = np.linspace(4.0e6, 0, 80)
  a, b = SO_ML(**conf), SO_ML(**conf)
  with pytest.raises(TypeError) as e:
    a.timestep(dt=dt, Psi_b=Psi_b)
  assert str(e.value) == 'b_basin needs to be numpy array providing buoyancy levels in basin'
  with pytest.raises(TypeError) as e:
    a.timestep(dt=dt, b_basin=b_basin)
  assert str(e.value) == 'Psi_b needs to be numpy array providing overturning at buoyancy levels given by b_basin'
  a.timestep(dt=dt, b_basin=b_basin, Psi_b=Psi_b)
  b.advdiff(b_basin=b_basin, Psi_b=Psi_b, dt=dt)
  assert all(a.bs == b.bs) and all(a.Psi_s == b.Psi_s)
  # the same call against the oracle (decreasing b_basin: numpy's guess-carrying search decides)
  from oracle import pymoc_oracle as O
  ref = O.MixedLayerState(**{k: (float(v) if not isinstance(v, np.ndarray) else v) for k, v in conf.items()})
  O.ml_timestep(ref, b_basin, Psi_b, dt)
  assert relmax(a.bs, ref.bs) < TOL and relmax(a.Psi_s, ref.Psi_s) < TOL, (relmax(a.bs, ref.bs), relmax(a.Psi_s, ref.Psi_s))
  # test_SO_ML.py:129-170 (5 % against the analytic tendency)
  y = np.asarray(np.linspace(0, 2.0e6, 51))
  Ks, L, h, surflux = 100, 4e6, 50, 5.9e3
  dtheta_dy = 2.0 * np.pi / 2.0e6
  b_basin = np.asarray([0.02 * (n / 2.0e6)**2 for n in y])
  bs = np.asarray([b_basin[-1] * np.cos(n * dtheta_dy) for n in y])
  Psi_b = np.asarray(np.linspace(1e4, 2.0e4, 51))
  ml = SO_ML(y=y, Ks=Ks, h=h, L=L, surflux=surflux, rest_mask=0.0, b_rest=0.0, v_pist=2.0 / 86400.0, bs=bs)
  dbs_dy = np.asarray([-dtheta_dy * b_basin[-1] * np.sin(n * dtheta_dy) for n in y])
  d2bs_dy2 = np.asarray([-dtheta_dy**2 * b_basin[-1] * np.cos(n * dtheta_dy) for n in y])
  db = -((Psi_b / (h * L)) * dbs_dy + Ks * d2bs_dy2 + surflux / h) * dt
  want = -(ml.bs.copy() + db)
  ml.advdiff(b_basin, Psi_b, dt)
  assert all(np.abs(want[i] - ml.bs[i]) / want[i] < 0.05 for i in range(len(want)))
  # fixtures produced by the reference's SO_ML.timestep
  for i, d in golden('units')['ml'].items():
    inp, out = d['inp'], d['out']
    ml = SO_ML(y=inp['y'], Ks=inp['Ks'], h=inp['h'], L=inp['L'], surflux=inp['surflux'].copy(),
               rest_mask=inp['rest_mask'].copy(), b_rest=inp['b_rest'].copy(), v_pist=inp['v_pist'], bs=inp['bs'].copy())
    for step in range(out['bs'].shape[0]):  # five successive steps
      ml.timestep(b_basin=inp['b_basin'], Psi_b=inp['Psi_b'], dt=inp['dt'])
      assert relmax(ml.Psi_s, out['Psi_s'][step]) < TOL, (i, step, relmax(ml.Psi_s, out['Psi_s'][step]))
      assert relmax(ml.bs, out['bs'][step]) < TOL, (i, step, relmax(ml.bs, out['bs'][step]))


# ---------------------------------------------------- isopycnal remap: fast and direct paths
def psib_edge_cases():
  """Psib / Psibz against the live oracle on profiles that steer the kernel through its O(nb+nz) path
  (non-decreasing columns, flat cells, flat cells sitting exactly on a class = the reference's 0/0)
  and its direct path (inverted cells), SURVEY H3."""
  from oracle import pymoc_oracle as O
  rng = np.random.default_rng(1)

  def check(tag, z, b1, b2, psi, nb=500, want_nan=None):
    with np.errstate(all='ignore'):
      want_p, want_g = O.thermwind_psib(psi, b1, b2, nb)
      wb, wn, _, _ = O.thermwind_psibz(psi, b1, b2, nb)
    tw = Psi_Thermwind(z=z, b1=b1.copy(), b2=b2.copy())
    tw.Psi = psi.copy()
    got = tw.Psib(nb)
    iso = tw.Psibz(nb)
    assert relmax(tw.bgrid, want_g) == 0.0, tag
    for g, w, key in ((got, want_p, 'psib'), (iso[0], wb, 'iso_b'), (iso[1], wn, 'iso_n')):
      assert relmax(g, w) < TOL, (tag, key, relmax(g, w))
    if want_nan is not None:
      assert int(np.isnan(want_p).sum()) == want_nan, (tag, int(np.isnan(want_p).sum()))

  for nz in (46, 80, 200):
    z = np.linspace(-4000, 0, nz)
    b1, b2 = 0.03 * np.exp(z / 300.), 0.004 * np.exp(z / 300.)
    psi = 10 * np.sin(np.pi * z / 4000.)**2 * np.sign(z + 1500)
    psi[0] = psi[-1] = 0
    check('monotone', z, b1, b2, psi)
    check('monotone nb=37', z, b1, b2, psi, 37)
    check('monotone nb=1000', z, b1, b2, psi, 1000)
    check('b2 flat zero', z, b1 - 0.001, 0 * z, psi)
    check('b2 flat on bgrid[0]', z, b1, 0 * z + b1.min(), psi, want_nan=1)
    c = b1.copy(); c[0] = c[1]
    check('flat bottom cell', z, c, b2, psi)
    c = b1.copy(); c[5:9] = c[5]
    check('flat run, down', z, c, b2, -np.abs(psi))
    check('flat run, up', z, c, b2, np.abs(psi))
    c = b2.copy(); c[0] = np.nextafter(c[1], 1.0)
    check('bottom cell inverted by one ulp', z, b1, c, psi)
    c = b1.copy(); c[10:20] = c[10:20][::-1]
    check('inverted region', z, c, b2, psi)
    pr = np.cumsum(rng.normal(size=nz)); pr[0] = 0
    check('random sorted', z, np.sort(rng.uniform(0, 0.03, nz)), np.sort(rng.uniform(0, 0.01, nz)), pr)
    ba = np.linspace(0, 0.03, nz)
    check('classes on the levels', z, ba, 0.5 * ba, pr, nb=nz)
    c = ba.copy(); c[7] = c[8]
    check('flat cell on a class', z, c, 0.5 * ba, pr, nb=nz)


# ------------------------------------------------------------------- host-side method surface
def host_api_checks():
  """Methods of the reference classes that do no time stepping (host arithmetic here as there)."""
  # tests/modules/test_psi_SO.py:135-142: linear b -> constant N2, 10 decimals
  z = np.asarray(np.linspace(-4000, 0, 80))
  y = np.asarray(np.linspace(0, 2.0e6, 51))
  so = Psi_SO(z=z, y=y, b=np.linspace(0.03, -0.001, 80), bs=0.05, tau=0.12)
  n2 = (so.b(z[1]) - so.b(z[0])) / (z[1] - z[0])
  f = so.calc_N2()
  assert all(np.round(f(zz), decimals=10) == np.round(n2, decimals=10) for zz in z)
  # non-uniform grid: the centred / one-sided differences of psi_SO.py:154-160
  zn = -4000. * (1. - np.linspace(0, 1, 41)**0.7)
  zn[-1] = 0.
  bn = 0.02 * np.exp(zn / 500.)
  so2 = Psi_SO(z=zn, y=y, b=bn, bs=0.05, tau=0.12)
  want = np.zeros(zn.size)
  h = zn[1:] - zn[:-1]
  want[1:-1] = (bn[2:] - bn[:-2]) / (h[1:] + h[:-1])
  want[0], want[-1] = (bn[1] - bn[0]) / h[0], (bn[-1] - bn[-2]) / h[-1]
  assert np.array_equal(so2.calc_N2()(zn), want)
  # tapers (psi_SO.py:164-216)
  assert so.calc_bottom_taper(None, z) == 1. and so.calc_top_taper(None, z) == 1.
  ek = so.calc_top_taper(None, z, scalar=False)
  assert ek.shape == z.shape and ek[-1] == 0. and np.all(ek[:-1] == 1.)
  bt = so.calc_bottom_taper(1000., z)
  assert bt[0] == 0. and np.all(bt[z >= z[0] + 1000.] == 1.) and np.all(np.diff(bt) >= 0)
  assert np.array_equal(bt, 1. - np.maximum(z[0] + 1000. - z, 0.)**2. / 1000.**2.)
  tt = so.calc_top_taper(500., z)
  assert tt[-1] == 0. and np.all(tt[z <= -500.] == 1.)
  # bc_GM (psi_SO.py:270-275)
  so.Psi_Ek = np.linspace(1., 2., 80)
  assert np.array_equal(so.bc_GM(np.array([3., 9.]), np.array([4., 9.])), np.array([3., 4.]))
  so.bvp_with_Ek = True
  assert np.array_equal(so.bc_GM(np.array([3., 9.]), np.array([4., 9.])), np.array([3. + 1e6, 4. + 2e6]))
  # Column.bc / ode / solve_equi against the reference's own runs (golden equi.npz, examples/example_iteration.py)
  t = golden('equi')
  zc = t['z']
  kappa = lambda zz: 1e-5 + 3e-5 * np.exp(zz / 100) + 3e-4 * np.exp(-zz / 1000 - 4)
  col = Column(z=zc, kappa=kappa, Area=float(t['A']), b=t['b0'].copy(), bs=float(t['bs']), bbot=float(t['bbot']))
  assert np.array_equal(col.bc(np.array([1., 2.]), np.array([3., 4.])), np.array([1. - col.bbot, 3. - col.bs]))
  alias = col.b
  for it in sorted(t['iters']):
    d = t['iters'][it]
    col.solve_equi(d['wA'])
    assert relmax(col.b, d['b']) < 1e-12 and relmax(col.bz, d['bz']) < 1e-12, it
  assert col.b is not alias  # rebinds, like the reference (column.py:207-208)
  col.bzbot = 1e-7
  assert np.array_equal(col.bc(np.array([1., 2.]), np.array([3., 4.])), np.array([2. - 1e-7, 3. - col.bs]))
  tw = Psi_Thermwind(z=zc, b1=t['b0'], b2=0.)
  assert np.array_equal(tw.bc(np.array([1., 2.]), np.array([3., 4.])), np.array([1., 3.]))
  assert np.array_equal(tw.ode(zc[:3], np.ones((2, 3)))[1], 1. / tw.f * (0. - t['b0'][:3]))
