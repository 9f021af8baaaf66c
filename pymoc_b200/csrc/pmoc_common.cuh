// Shared declarations of the pymoc_b200 translation units (kernels are split per
// levels-per-lane so that the instantiations compile in parallel).
#pragma once
#include "pmoc_device.cuh"
#include "../../include/pymoc_b200.h"

#include <cstdio>
#include <cstring>

extern thread_local char pmoc_g_err[256];
#define g_err pmoc_g_err

namespace pmk {



#ifndef PMOC_EMU
#define PM_CUDA_OK(expr)                                                                      \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      std::snprintf(g_err, sizeof(g_err), "%s: %s", #expr, cudaGetErrorString(e_));           \
      return e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? PMOC_ENODEVICE : PMOC_ECUDA; \
    }                                                                                         \
  } while (0)
#endif

static inline int fail(int code, const char* msg) {
  std::snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

PM_DEV double vat(const pmoc_vec& v, long long m) { return v.ptr[m * v.mstride]; }
PM_DEV const double* vrow(const pmoc_vec& v, long long m) { return v.ptr + m * v.mstride; }

constexpr int kWarpsPerBlock = 4;

// ------------------------------------------------------------------------------------------
// shared-memory plan of k_model (in doubles)
struct SmemPlan {
  int nzp, nyp, nbp;
  int off_z, off_y;        // per block
  int off_warp0, per_warp; // per warp region
  int w_ctop, w_crinv, w_cu, w_psib, w_bs;
  size_t bytes(int wpb) const { return sizeof(double) * (size_t)(off_warp0 + per_warp * wpb); }
};

static inline SmemPlan plan_smem(int LPL, int ny, int nb, unsigned flags) {
  SmemPlan s{};
  s.nzp = 32 * LPL;
  s.nyp = (ny + 3) & ~3;
  s.nbp = (nb + 3) & ~3;
  int o = 0;
  s.off_z = o; o += s.nzp + 4;
  s.off_y = o; o += s.nyp;
  s.off_warp0 = o;
  int w = 0;
  if (flags & PMOC_ISO) {
    s.w_ctop = w; w += s.nzp;
    s.w_crinv = w; w += s.nzp;
    s.w_cu = w; w += s.nzp;
    s.w_psib = w; w += s.nbp;
  }
  if (flags & PMOC_HAS_SO) { s.w_bs = w; w += s.nyp; }
  s.per_warp = w;
  return s;
}

struct RunArgs {
  pmoc_model m;
  SmemPlan sp;
  long long it0, nsteps;
  int diagnose_only;
};

// per-column registers
template <int LPL>
struct ColRegs {
  double b[LPL], p[LPL], q[LPL];
  double bs, N2min, bbot, bzbot;
  bool has_bzbot, conv;
  int var;
};

template <int LPL>
PM_DEV void col_load(ColRegs<LPL>& c, const pmoc_column& d, long long m, int nz) {
  pm::load_lev<LPL>(c.b, d.b + m * nz, nz, 0.0);
  c.bs = vat(d.bs, m);
  c.N2min = vat(d.N2min, m);
  c.bbot = d.bbot[m];
  c.has_bzbot = d.bzbot.ptr != nullptr;
  c.bzbot = c.has_bzbot ? vat(d.bzbot, m) : 0.0;
  c.conv = d.do_conv != 0;
  c.var = (d.var != nullptr && d.nvar > 1) ? d.var[m] : 0;
}

template <int LPL>
PM_DEV void col_refold(ColRegs<LPL>& c, const pmoc_column& d, long long m, const double (&wA)[LPL],
                       const double* zs, int nz, double dt) {
  const long long voff = (long long)c.var * nz;
  pm::col_coeffs<LPL>(c.p, c.q, wA, vrow(d.kappa, m) + voff, vrow(d.dAk, m) + voff, vrow(d.Area, m), zs, nz, dt);
}

// Column.timestep(wA, dt, do_conv) as the loop calls it (column.py:336-341)
template <int LPL>
PM_DEV void col_advance(ColRegs<LPL>& c, const double* zs, int nz) {
  if (c.conv) pm::col_convect<LPL>(c.b, c.bs, c.N2min, zs, nz);
  if (rt::lane() == 0) {
    if (c.has_bzbot) {
      if (LPL > 1) c.b[0] = c.b[LPL > 1 ? 1 : 0] - c.bzbot * (zs[1] - zs[0]);
    } else {
      c.b[0] = c.bbot;
    }
  }
  pm::col_step<LPL>(c.b, c.p, c.q);
}

// ------------------------------------------------------------------------------------------
// launch plumbing
#ifdef PMOC_EMU
template <class K, class A>
int launch(K kern, long long grid, int block, size_t smem, void*, const A& args) {
  pmemu::launch(grid, block, smem, [=]() { kern(args); });
  return PMOC_OK;
}
#else
template <class K, class A>
int launch(K kern, long long grid, int block, size_t smem, void* stream, const A& args) {
  if (smem > 48 * 1024)
    PM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)grid, block, smem, (cudaStream_t)stream>>>(args);
  PM_CUDA_OK(cudaGetLastError());
  return PMOC_OK;
}
#endif


static inline long long blocks_for(long long M) { return (M + kWarpsPerBlock - 1) / kWarpsPerBlock; }
}  // namespace pmk
using namespace pmk;

// one launcher per levels-per-lane, each in its own translation unit (pmoc_model.cu)
#define PM_DECL_MODEL(n) int pmoc_launch_model_##n(const RunArgs& ra, void* stream);
PM_DECL_MODEL(2) PM_DECL_MODEL(3) PM_DECL_MODEL(4) PM_DECL_MODEL(5) PM_DECL_MODEL(6) PM_DECL_MODEL(7) PM_DECL_MODEL(8)
