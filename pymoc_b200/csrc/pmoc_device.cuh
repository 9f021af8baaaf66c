// Warp-level building blocks of the fused PyMOC engine (fp64 throughout).
//
// Layout: one warp owns one ensemble member.  A vertical profile of nz <= 32*LPL levels is
// distributed in contiguous chunks -- lane L holds levels [L*LPL, (L+1)*LPL) in registers --
// so the explicit column step needs only two neighbour shuffles per step and everything
// else is lane-local.  All columns of a member share this map, hence the thermal-wind and
// isopycnal-remap arithmetic that combines basin and north values is lane-local too.
//
// Arithmetic policy (see pmoc_rt.cuh): the once-per-K-steps diagnostics follow the
// reference's NumPy expressions operation by operation (IEEE divide, no contraction);
// the per-step column update uses folded stencil coefficients (three FP64-pipe
// instructions per level and step).  file:line citations are under
// /root/reference/src/pymoc/modules.
#pragma once
#include "pmoc_rt.cuh"

namespace pm {

constexpr double kSv = 1e6;

template <int LPL>
PM_DEV int lev(int j) { return rt::lane() * LPL + j; }
// per-warp shared arrays that are only ever read by their owner use the conflict-free
// lane-major layout: slot j of lane L lives at j*32 + L
PM_DEV int lm(int j) { return j * 32 + rt::lane(); }

// x / c for a constant c with rc = 1/c pre-rounded: product, exact residual, one correction
// (Markstein).  Correctly rounded -- the same bits as the IEEE divide the reference performs --
// whenever no intermediate over/underflows; checked against x / 1e6 on 4e8 random doubles.
PM_DEV double div_const(double x, double c, double rc) {
  const double q = x * rc;
  return rt::fma(rt::fma(-c, q, x), rc, q);
}

// registers <- natural-order global/shared array
template <int LPL>
PM_DEV void load_lev(double (&v)[LPL], const double* PM_RESTRICT g, int n, double pad) {
  PM_UNROLL
  for (int j = 0; j < LPL; ++j) {
    const int i = lev<LPL>(j);
    v[j] = i < n ? g[i] : pad;
  }
}
template <int LPL>
PM_DEV void store_lev(const double (&v)[LPL], double* PM_RESTRICT g, int n) {
  PM_UNROLL
  for (int j = 0; j < LPL; ++j) {
    const int i = lev<LPL>(j);
    if (i < n) g[i] = v[j];
  }
}

// exclusive prefix sum of one value per lane
PM_DEV double wscan_excl(double v) {
  const int L = rt::lane();
  PM_UNROLL
  for (int d = 1; d < 32; d <<= 1) {
    const double t = rt::shfl_up(v, d);
    if (L >= d) v = v + t;
  }
  const double prev = rt::shfl_up(v, 1);
  return L == 0 ? 0.0 : prev;
}

// ===================================================================== Column
// Folded coefficients of the explicit step (column.py:235-249).  With d_i = b[i+1]-b[i]:
//   b_i += dt*( -weff_i * (weff_i<0 ? d_i/dzu : d_{i-1}/dzd) / A_i
//               + kappa_i * (d_i/dzu - d_{i-1}/dzd) / (0.5*(dzu+dzd)) )  =  p_i d_i - q_i d_{i-1}
// weff = wA - d(A kappa)/dz is fixed between two streamfunction updates, so p, q are rebuilt
// only then.  Everything state-independent is tabulated once per launch:
//   per block  (GeoTab, lane-major): 1/dzu, 1/dzd, 1/(dzc dzu), 1/(dzc dzd), z
//   per member (ColTab, lane-major): KU = dt kappa/(dzc dzu), KD = dt kappa/(dzc dzd), RA = dt/A, dAk
// Boundary and padding levels get p = q = 0.
struct GeoTab {
  const double *zs;                      // natural order, nzp+4 (padded with z[nz-1])
  const double *zl, *rdu, *rdd, *ruu, *rdd2;  // lane-major, nzp each
};
struct ColTab {
  double *ku, *kd, *ra, *dak;  // lane-major, nzp each (per warp)
};

// block-cooperative fill of the geometry tables (call before the block barrier)
template <int LPL>
PM_DEV void geo_fill(double* zs, double* zl, double* rdu, double* rdd, double* ruu, double* rdd2,
                     const double* PM_RESTRICT z, int nz, int tid, int nthr) {
  const int nzp = 32 * LPL;
  for (int i = tid; i < nzp + 4; i += nthr) zs[i] = z[i < nz ? i : nz - 1];
  for (int s = tid; s < nzp; s += nthr) {
    const int L = s & 31, j = s >> 5, i = L * LPL + j;
    double a = 0., b = 0., c = 0., d = 0.;
    if (i >= 1 && i < nz - 1) {
      const double dzu = z[i + 1] - z[i], dzd = z[i] - z[i - 1], dzc = 0.5 * (dzu + dzd);
      a = 1. / dzu;
      b = 1. / dzd;
      c = 1. / (dzc * dzu);
      d = 1. / (dzc * dzd);
    }
    zl[s] = z[i < nz ? i : nz - 1];
    rdu[s] = a;
    rdd[s] = b;
    ruu[s] = c;
    rdd2[s] = d;
  }
}

// per-member tables from the (host-sampled) kappa, d(A kappa)/dz and Area profiles
template <int LPL>
PM_DEV void col_tabulate(const ColTab& T, const GeoTab& G, const double* PM_RESTRICT kappa,
                         const double* PM_RESTRICT dAk, const double* PM_RESTRICT Area, int nz, double dt) {
  PM_UNROLL
  for (int j = 0; j < LPL; ++j) {
    const int i = lev<LPL>(j), s = lm(j);
    double ku = 0., kd = 0., ra = 0., dk = 0.;
    if (i >= 1 && i < nz - 1) {
      const double kdt = dt * kappa[i];
      ku = kdt * G.ruu[s];
      kd = kdt * G.rdd2[s];
      ra = dt / Area[i];
      dk = dAk[i];
    }
    T.ku[s] = ku;
    T.kd[s] = kd;
    T.ra[s] = ra;
    T.dak[s] = dk;
  }
  rt::syncwarp();
}

template <int LPL>
PM_DEV void col_coeffs(double (&p)[LPL], double (&q)[LPL], const double (&wA)[LPL], const ColTab& T,
                       const GeoTab& G, int nz) {
  PM_UNROLL
  for (int j = 0; j < LPL; ++j) {
    const int i = lev<LPL>(j), s = lm(j);
    double pj = 0.0, qj = 0.0;
    if (i >= 1 && i < nz - 1) {
      const double weff = wA[j] - T.dak[s];
      const double ra = T.ra[s];
      pj = T.ku[s];
      qj = T.kd[s];
      if (weff < 0)
        pj = pj - weff * (ra * G.rdu[s]);
      else
        qj = qj + weff * (ra * G.rdd[s]);
    }
    p[j] = pj;
    q[j] = qj;
  }
}

// One explicit step: 2 neighbour shuffles, LPL+1 subtractions, 2*LPL FMAs.
template <int LPL>
PM_DEV void col_step(double (&b)[LPL], const double (&p)[LPL], const double (&q)[LPL]) {
  const double bnext = rt::shfl_down(b[0], 1);
  const double bprev = rt::shfl_up(b[LPL - 1], 1);
  double dm = b[0] - bprev;
  PM_UNROLL
  for (int j = 0; j < LPL; ++j) {
    const double d = (j < LPL - 1 ? b[j + 1 < LPL ? j + 1 : j] : bnext) - b[j];
    b[j] = rt::fma(-q[j], dm, rt::fma(p[j], d, b[j]));
    dm = d;
  }
}

// Convective adjustment (column.py:264-271).  Strict '>' and un-fused bs + N2min*(z - zconv).
template <int LPL>
PM_DEV void col_convect(double (&b)[LPL], double bs, double N2min, const double* zs, const double* zl, int nz) {
  unsigned mine = 0;
  int top_stable = -1;
  PM_UNROLL
  for (int j = 0; j < LPL; ++j) {
    const int i = lev<LPL>(j);
    if (i < nz) {
      if (b[j] > bs)
        mine |= 1u << j;
      else
        top_stable = i;
    }
  }
  if (rt::ballot(mine != 0)) {
    const int anchor = rt::max_i(top_stable);
    const double zc = zs[anchor >= 0 ? anchor : 0];
    PM_UNROLL
    for (int j = 0; j < LPL; ++j)
      if ((mine >> j) & 1u) b[j] = bs + N2min * (zl[lm(j)] - zc);
  } else {
    PM_UNROLL
    for (int j = 0; j < LPL; ++j)
      if (lev<LPL>(j) == nz - 1) b[j] = bs;
  }
}

// Lateral inflow (column.py:306-313)
template <int LPL>
PM_DEV void col_horadv(double (&b)[LPL], const double* PM_RESTRICT vdx, const double* PM_RESTRICT b_in,
                       const double* PM_RESTRICT Area, int nz, double dt) {
  PM_UNROLL
  for (int j = 0; j < LPL; ++j) {
    const int i = lev<LPL>(j);
    if (i < nz && vdx[i] > 0.0) b[j] = b[j] + dt * vdx[i] * (b_in[i] - b[j]) / Area[i];
  }
}

template <int LPL>
PM_DEV void set_level(double (&b)[LPL], int level, double v) {
  PM_UNROLL
  for (int j = 0; j < LPL; ++j)
    if (lev<LPL>(j) == level) b[j] = v;
}
// value of one level, broadcast to the warp
template <int LPL>
PM_DEV double get_level(const double (&b)[LPL], int level) {
  double mine = 0.0;
  PM_UNROLL
  for (int j = 0; j < LPL; ++j)
    if (lev<LPL>(j) == level) mine = b[j];
  return rt::shfl(mine, level / LPL);
}

// ============================================================== Psi_Thermwind
// Exact solution of Psi'' = (b2-b1)/f, Psi(z0)=Psi(zN)=0 (psi_thermwind.py:123-135) in the form
// SciPy's collocation produces on the un-refined mesh z: Simpson for Psi' with the mid-point
// value gm, cubic-Hermite Simpson for Psi.  For piecewise-linear b's (gm = mean of the ends)
// it is the exact C1 piecewise cubic.  Result in Sv.
template <int LPL>
PM_DEV void tw_solve(double (&psi)[LPL], const double (&b1)[LPL], const double (&b2)[LPL], double f,
                     const double* zs, int nz, const double* PM_RESTRICT gmid) {
  const double rf = 1. / f;
  double g[LPL], T[LPL], part[LPL];
  PM_UNROLL
  for (int j = 0; j < LPL; ++j) g[j] = lev<LPL>(j) < nz ? rf * (b2[j] - b1[j]) : 0.0;
  const double gnext = rt::shfl_down(g[0], 1);
  double run = 0.0;
  PM_UNROLL
  for (int j = 0; j < LPL; ++j) {
    const int i = lev<LPL>(j);
    const double gi = g[j], gi1 = j < LPL - 1 ? g[j + 1 < LPL ? j + 1 : j] : gnext;
    double t = 0.0;
    if (i < nz - 1) {
      const double h = zs[i + 1] - zs[i];
      const double gm = gmid ? rf * gmid[i] : 0.5 * (gi + gi1);
      t = h / 6. * (gi + 4. * gm + gi1);
    }
    T[j] = t;
    part[j] = run;
    run = run + t;
  }
  const double base1 = wscan_excl(run);
  run = 0.0;
  PM_UNROLL
  for (int j = 0; j < LPL; ++j) {
    const int i = lev<LPL>(j);
    const double gi = g[j], gi1 = j < LPL - 1 ? g[j + 1 < LPL ? j + 1 : j] : gnext;
    double cell = 0.0;
    if (i < nz - 1) {
      const double h = zs[i + 1] - zs[i];
      cell = h * (base1 + part[j]) + h * T[j] / 2. - h * h / 12. * (gi1 - gi);
    }
    part[j] = run;
    run = run + cell;
  }
  const double base2 = wscan_excl(run);
  PM_UNROLL
  for (int j = 0; j < LPL; ++j) part[j] = base2 + part[j];
  const double total = get_level<LPL>(part, nz - 1);
  const double z0 = zs[0], H = zs[nz - 1] - zs[0];
  PM_UNROLL
  for (int j = 0; j < LPL; ++j) {
    const int i = lev<LPL>(j);
    psi[j] = i < nz ? div_const(part[j] - total * ((zs[i] - z0) / H), kSv, 1.0 / kSv) : 0.0;
  }
}

// np.linspace(bmin, bmax, nb)[i]  (numpy: arange(nb)*step + start, last element = stop)
struct BGrid {
  double lo, hi, step;
  int nb;
  PM_DEV double at(int i) const { return i == nb - 1 ? hi : (double)i * step + lo; }
};

// Upwind isopycnal remap (psi_thermwind.py:170-185).  Cell data go to shared memory
// (ctop/crinv/cu, natural order), every lane then owns KB classes of bgrid per pass and sweeps
// all cells:  psib[i] = sum_c clip((top_c - bgrid_i)/(top_c - bot_c), 0, 1) * u_c.
// The divide is by a per-cell reciprocal (same +-inf / NaN outcomes for flat cells);
// the comparisons-based clip keeps NaN like np.clip.
template <int LPL>
PM_DEV BGrid tw_psib(const double (&psi)[LPL], const double (&b1)[LPL], const double (&b2)[LPL], int nz, int nb,
                     double* ctop, double* crinv, double* cu, double* psib_s) {
  const double psin = rt::shfl_down(psi[0], 1);
  const double b1n = rt::shfl_down(b1[0], 1), b2n = rt::shfl_down(b2[0], 1);
  double lo = INFINITY, hi = -INFINITY;
  PM_UNROLL
  for (int j = 0; j < LPL; ++j) {
    const int i = lev<LPL>(j);
    if (i < nz) {
      lo = b1[j] < lo ? b1[j] : lo;
      lo = b2[j] < lo ? b2[j] : lo;
      hi = b1[j] > hi ? b1[j] : hi;
      hi = b2[j] > hi ? b2[j] : hi;
    }
    if (i < nz - 1) {
      const bool last = j == LPL - 1;
      const int jn = j + 1 < LPL ? j + 1 : j;
      const double u = -((last ? psin : psi[jn]) - psi[j]);
      const bool from2 = u < 0;
      const double bot = from2 ? b2[j] : b1[j];
      const double top = from2 ? (last ? b2n : b2[jn]) : (last ? b1n : b1[jn]);
      ctop[i] = top;
      crinv[i] = 1.0 / (top - bot);
      cu[i] = u;
    }
  }
  BGrid G;
  G.lo = rt::wmin(lo);
  G.hi = rt::wmax(hi);
  G.nb = nb;
  G.step = (G.hi - G.lo) / (double)(nb - 1);
  rt::syncwarp();
  constexpr int KB = 8;
  const int L = rt::lane();
  for (int base = 0; base < nb; base += 32 * KB) {
    double bg[KB], acc[KB];
    PM_UNROLL
    for (int k = 0; k < KB; ++k) {
      const int i = base + k * 32 + L;
      bg[k] = G.at(i < nb ? i : nb - 1);
      acc[k] = 0.0;
    }
    for (int c = 0; c < nz - 1; ++c) {
      const double top = ctop[c], r = crinv[c], u = cu[c];
      PM_UNROLL
      for (int k = 0; k < KB; ++k) {
        double t = (top - bg[k]) * r;
        t = t < 0. ? 0. : t;
        t = t > 1. ? 1. : t;
        acc[k] = rt::fma(t, u, acc[k]);
      }
    }
    PM_UNROLL
    for (int k = 0; k < KB; ++k) {
      const int i = base + k * 32 + L;
      if (i < nb) psib_s[i] = acc[k];
    }
  }
  rt::syncwarp();
  return G;
}

// np.interp(x, bgrid, psib) (numpy/_core/src/multiarray/compiled_base.c: arr_interp) on the
// implicit uniform grid: index by division, then corrected against the actual grid values.
PM_DEV double interp_bgrid(double x, const BGrid& G, const double* psib_s) {
  if (x != x) return x;
  const int nb = G.nb;
  if (x > G.hi) return psib_s[nb - 1];
  if (x < G.lo) return psib_s[0];
  int j = nb - 1;
  if (G.step > 0.) {
    const double r = (x - G.lo) / G.step;
    j = r < (double)(nb - 1) ? (int)r : nb - 1;
    while (j < nb - 1 && G.at(j + 1) <= x) ++j;
    while (j > 0 && G.at(j) > x) --j;
  }
  if (j == nb - 1) return psib_s[j];
  const double xj = G.at(j);
  const double fj = psib_s[j];
  if (xj == x) return fj;
  const double xj1 = G.at(j + 1), fj1 = psib_s[j + 1];
  const double slope = (fj1 - fj) / (xj1 - xj);
  double res = slope * (x - xj) + fj;
  if (res != res) {
    res = slope * (x - xj1) + fj1;
    if (res != res && fj == fj1) res = fj;
  }
  return res;
}

// ===================================================================== Psi_SO
// index of the last element <= key in a non-decreasing array (numpy binary_search_with_guess
// result for a key inside the range)
PM_DEV int search_le(const double* a, int n, double key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);
    if (key >= a[mid])
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo - 1;
}

// np.interp(x, xp, fp) for a scalar x, generic (possibly non-uniform) increasing xp
PM_DEV double interp1(double x, const double* xp, const double* fp, int n) {
  if (x != x) return x;
  if (x > xp[n - 1]) return fp[n - 1];
  if (x < xp[0]) return fp[0];
  const int j = search_le(xp, n, x);
  if (j == n - 1) return fp[j];
  if (xp[j] == x) return fp[j];
  const double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
  double res = slope * (x - xp[j]) + fp[j];
  if (res != res) {
    res = slope * (x - xp[j + 1]) + fp[j + 1];
    if (res != res && fp[j] == fp[j + 1]) res = fp[j];
  }
  return res;
}

PM_DEV bool sgn(double v) { return std::signbit(v); }

// scipy.optimize.brentq(lambda y: bs(y) - bval, ya, yb) with the default tolerances --
// statement-for-statement the iteration of scipy/optimize/Zeros/brentq.c (Brent 1973), so a
// multi-root bs(y) resolves to the root the reference finds.
PM_DEV double outcrop_brent(double bval, const double* ygrid, const double* bs, int ny, int south, bool* sign_error) {
  const double xtol = 2e-12, rtol = 8.881784197001252e-16;
  double xpre = ygrid[south], xcur = ygrid[ny - 1];
  double xblk = 0., fblk = 0., spre = 0., scur = 0.;
  double fpre = interp1(xpre, ygrid, bs, ny) - bval;
  double fcur = interp1(xcur, ygrid, bs, ny) - bval;
  if (fpre == 0) return xpre;
  if (fcur == 0) return xcur;
  if (sgn(fpre) == sgn(fcur)) {
    *sign_error = true;
    return NAN;
  }
  for (int it = 0; it < 100; ++it) {
    if (fpre != 0 && fcur != 0 && sgn(fpre) != sgn(fcur)) {
      xblk = xpre;
      fblk = fpre;
      spre = scur = xcur - xpre;
    }
    if (fabs(fblk) < fabs(fcur)) {
      xpre = xcur; xcur = xblk; xblk = xpre;
      fpre = fcur; fcur = fblk; fblk = fpre;
    }
    const double delta = (xtol + rtol * fabs(xcur)) / 2;
    const double sbis = (xblk - xcur) / 2;
    if (fcur == 0 || fabs(sbis) < delta) return xcur;
    if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
      double stry;
      if (xpre == xblk) {
        stry = -fcur * (xcur - xpre) / (fcur - fpre);
      } else {
        const double dpre = (fpre - fcur) / (xpre - xcur);
        const double dblk = (fblk - fcur) / (xblk - xcur);
        stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
      }
      const double lim1 = fabs(spre), lim2 = 3 * fabs(sbis) - delta;
      if (2 * fabs(stry) < (lim1 < lim2 ? lim1 : lim2)) {
        spre = scur;
        scur = stry;
      } else {
        spre = sbis;
        scur = sbis;
      }
    } else {
      spre = sbis;
      scur = sbis;
    }
    xpre = xcur;
    fpre = fcur;
    if (fabs(scur) > delta)
      xcur += scur;
    else
      xcur += (sbis > 0 ? delta : -delta);
    fcur = interp1(xcur, ygrid, bs, ny) - bval;
  }
  return xcur;
}

// Unique root of the piecewise-linear bs(y) = bval when bs is non-decreasing north of its
// minimum.  sinv[k] = (y[k+1]-y[k])/(bs[k+1]-bs[k]) is tabulated with the surface scan.
PM_DEV double outcrop_monotone(double bval, const double* ygrid, const double* bs, const double* sinv, int ny,
                               int south) {
  if (bs[south] == bval) return ygrid[south];
  if (bs[ny - 1] == bval) return ygrid[ny - 1];
  int lo = south, hi = ny - 1;  // first index with bs >= bval lies in (south, ny-1]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (bs[mid] >= bval)
      hi = mid;
    else
      lo = mid;
  }
  if (bs[hi] == bval) return ygrid[hi];
  return ygrid[hi - 1] + (bval - bs[hi - 1]) * sinv[hi - 1];
}

// np.mean(tau + 0*np.linspace(y0, yN, 100)) for a float tau (psi_SO.py:239): numpy's pairwise
// reduction of 100 equal terms (8 accumulators over 96 terms, tree combine, 4 trailing adds).
PM_DEV double mean100(double tau) {
  double r = tau;
  for (int k = 0; k < 11; ++k) r = r + tau;
  double res = ((r + r) + (r + r)) + ((r + r) + (r + r));
  for (int k = 0; k < 4; ++k) res = res + tau;
  return res / 100.0;
}

struct SoPar {
  double tau_ave, f, rho, L, KGM, smax;
  const double *sill, *ektap, *toptap, *bottap;  // [nz] natural order, global
};
struct SoSurf {  // per-refresh scan of bs(y)
  double mn, bsN, y0, yN;
  int south;
  bool mono;
};
// Scan of the surface buoyancy: minimum / argmin (first occurrence, np.argmin), monotonicity
// north of it, and the inverse segment slopes.  Warp-cooperative; redone only when bs changes.
PM_DEV SoSurf so_scan(const double* ygrid, const double* bs, double* sinv, int ny) {
  SoSurf s;
  s.mn = bs[0];
  s.south = 0;
  for (int k = 1; k < ny; ++k)
    if (bs[k] < s.mn) {
      s.mn = bs[k];
      s.south = k;
    }
  bool down = false;
  for (int k = rt::lane(); k < ny - 1; k += 32) {
    const double db = bs[k + 1] - bs[k];
    sinv[k] = (ygrid[k + 1] - ygrid[k]) / db;
    if (k >= s.south && db < 0) down = true;
  }
  s.mono = rt::ballot(down) == 0;
  s.bsN = bs[ny - 1];
  s.y0 = ygrid[0];
  s.yN = ygrid[ny - 1];
  rt::syncwarp();
  return s;
}

// Psi_SO.solve with the explicit GM branch (psi_SO.py:106-140, 218-243, 302-354).  Sv.
template <int LPL>
PM_DEV void so_solve(double (&psi)[LPL], double (&ek)[LPL], double (&gm)[LPL], double (&ysv)[LPL],
                     const double (&b)[LPL], const double* ygrid, const double* bs, const double* sinv, int ny,
                     const SoSurf& S, const SoPar& P, const double* zs, int nz, unsigned* status) {
  if (!S.mono) *status |= 2u;
  const double pre = P.tau_ave / P.f / P.rho * P.L;
  const double c6 = 1e6, r6 = 1.0 / 1e6;
  PM_UNROLL
  for (int j = 0; j < LPL; ++j) {
    const int i = lev<LPL>(j);
    double e = 0., g = 0., ps = 0., yo = 0.;
    if (i < nz) {
      const double bi = b[j];
      if (bi < S.mn)
        yo = S.y0 - 1e3;
      else if (bi > S.bsN)
        yo = S.yN;
      else if (S.mono)
        yo = outcrop_monotone(bi, ygrid, bs, sinv, ny, S.south);
      else {
        bool bad = false;
        yo = outcrop_brent(bi, ygrid, bs, ny, S.south, &bad);
        if (bad) *status |= 4u;
      }
      e = div_const(pre * P.sill[i] * P.ektap[i], c6, r6);
      double dy = S.yN - yo;
      dy = 0.1 > dy ? 0.1 : dy;
      const double s = zs[i] / dy, ms = -P.smax;
      const double mx = (s >= ms || s != s) ? s : ms;
      double t = P.KGM * mx * P.L * P.toptap[i] * P.bottap[i];
      if (dy > S.yN - S.y0) {
        const double alt = -e * 1e6;
        t = (t >= alt || t != t) ? t : alt;
      }
      g = div_const(t, c6, r6);
      ps = i == 0 ? 0. : e + g;
    }
    ek[j] = e;
    gm[j] = g;
    psi[j] = ps;
    ysv[j] = yo;
  }
}

}  // namespace pm
