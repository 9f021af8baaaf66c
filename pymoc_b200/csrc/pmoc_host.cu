// Host-buffer entry point (pmoc_model_run_host) and the FP64 roofline probe.
#include "pmoc_common.cuh"

#include <vector>

#ifndef PMOC_EMU
namespace {

// Mirrors host arrays of a pmoc_model on the device for the duration of one call.
struct Mirror {
  cudaStream_t s = nullptr;
  struct Back { void* host; void* dev; size_t bytes; };
  std::vector<void*> allocs;
  std::vector<Back> backs;
  cudaError_t err = cudaSuccess;

  void* alloc(size_t bytes) {
    void* d = nullptr;
    if (err == cudaSuccess) err = cudaMallocAsync(&d, bytes ? bytes : 8, s);
    if (d) allocs.push_back(d);
    return d;
  }
  // input: upload
  template <class T>
  const T* in(const T* h, size_t n) {
    if (!h) return nullptr;
    void* d = alloc(n * sizeof(T));
    if (d && err == cudaSuccess) err = cudaMemcpyAsync(d, h, n * sizeof(T), cudaMemcpyHostToDevice, s);
    return static_cast<const T*>(d);
  }
  pmoc_vec in(pmoc_vec v, long long M, size_t len) {
    if (!v.ptr) return v;
    const size_t n = v.mstride ? (size_t)(M - 1) * (size_t)v.mstride + len : len;
    return pmoc_vec{in(v.ptr, n), v.mstride};
  }
  // output (optionally also an input): copied back at the end
  template <class T>
  T* out(T* h, size_t n, bool upload) {
    if (!h) return nullptr;
    void* d = alloc(n * sizeof(T));
    if (d && err == cudaSuccess) {
      err = upload ? cudaMemcpyAsync(d, h, n * sizeof(T), cudaMemcpyHostToDevice, s)
                   : cudaMemsetAsync(d, 0, n * sizeof(T), s);
      backs.push_back({h, d, n * sizeof(T)});
    }
    return static_cast<T*>(d);
  }
  void copy_back() {
    for (auto& b : backs)
      if (err == cudaSuccess) err = cudaMemcpyAsync(b.host, b.dev, b.bytes, cudaMemcpyDeviceToHost, s);
  }
  void release() {
    for (void* d : allocs) cudaFreeAsync(d, s);
    allocs.clear();
  }
};

pmoc_column mirror_column(Mirror& mr, const pmoc_column& c, long long M, int nz) {
  pmoc_column d = c;
  d.b = mr.out(c.b, (size_t)M * nz, true);
  d.kappa = mr.in(c.kappa, M, (size_t)c.nvar * nz);
  d.dAk = mr.in(c.dAk, M, (size_t)c.nvar * nz);
  d.Area = mr.in(c.Area, M, nz);
  d.bs = mr.in(c.bs, M, 1);
  d.N2min = mr.in(c.N2min, M, 1);
  d.bzbot = mr.in(c.bzbot, M, 1);
  d.bbot = mr.out(c.bbot, (size_t)M, true);
  d.var = mr.out(c.var, (size_t)M, true);
  return d;
}

}  // namespace
#endif

extern "C" int pmoc_model_run_host(const pmoc_model* m, int64_t it0, int64_t nsteps) {
  if (!m) return fail(PMOC_EINVAL, "model is NULL");
#ifdef PMOC_EMU
  if (it0 == 0 && !(m->flags & PMOC_ORDER_JN))
    if (int rc = pmoc_model_diagnose(m, nullptr)) return rc;
  return pmoc_model_run(m, it0, nsteps, nullptr);
#else
  if (m->M <= 0 || m->nz < 3) return fail(PMOC_EINVAL, "bad M / nz");
  const long long M = m->M;
  const int nz = m->nz, ny = m->ny, nb = m->nb;
  const unsigned f = m->flags;
  Mirror mr;
  {  // keep the stream-ordered pool's memory between calls (the default trims it at every synchronisation)
    int dev = 0;
    cudaMemPool_t pool;
    unsigned long long keep = ~0ull;
    PM_CUDA_OK(cudaGetDevice(&dev));
    PM_CUDA_OK(cudaDeviceGetDefaultMemPool(&pool, dev));
    PM_CUDA_OK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  PM_CUDA_OK(cudaStreamCreateWithFlags(&mr.s, cudaStreamNonBlocking));
  pmoc_model d = *m;
  const bool carry = it0 > 0;  // streamfunctions diagnosed by an earlier call are inputs
  d.z = mr.in(m->z, nz);
  d.y = mr.in(m->y, ny);
  d.basin = mirror_column(mr, m->basin, M, nz);
  if (f & PMOC_HAS_NORTH) d.north = mirror_column(mr, m->north, M, nz);
  if (f & PMOC_HAS_PAC) {
    d.pac = mirror_column(mr, m->pac, M, nz);
    d.zoc_f = mr.in(m->zoc_f, M, 1);
    d.so2_L = mr.in(m->so2_L, M, 1);
    d.Psi_zoc = mr.out(m->Psi_zoc, (size_t)M * nz, carry);
    d.Psi_zon_a = mr.out(m->Psi_zon_a, (size_t)M * nz, carry);
    d.Psi_zon_p = mr.out(m->Psi_zon_p, (size_t)M * nz, carry);
    d.psib2 = mr.out(m->psib2, (size_t)M * nb, carry);
    d.bgrid2 = mr.out(m->bgrid2, (size_t)M * nb, carry);
    d.Psi_so2 = mr.out(m->Psi_so2, (size_t)M * nz, carry);
    d.Psi_Ek2 = mr.out(m->Psi_Ek2, (size_t)M * nz, carry);
    d.Psi_GM2 = mr.out(m->Psi_GM2, (size_t)M * nz, carry);
  }
  d.tw_f = mr.in(m->tw_f, M, 1);
  d.tw_b2 = mr.in(m->tw_b2, M, nz);
  d.so_bs = mr.in(m->so_bs, M, ny);
  d.so_tau = mr.in(m->so_tau, M, m->so_tau_on_y ? ny : 1);
  d.so_f = mr.in(m->so_f, M, 1);
  d.so_rho = mr.in(m->so_rho, M, 1);
  d.so_L = mr.in(m->so_L, M, 1);
  d.so_KGM = mr.in(m->so_KGM, M, 1);
  d.so_smax = mr.in(m->so_smax, M, 1);
  d.so_c = mr.in(m->so_c, M, 1);
  d.so_sill_taper = mr.in(m->so_sill_taper, nz);
  d.so_ek_taper = mr.in(m->so_ek_taper, nz);
  d.so_top_taper = mr.in(m->so_top_taper, nz);
  d.so_bot_taper = mr.in(m->so_bot_taper, nz);
  d.ml_bs = mr.out(m->ml_bs, (size_t)M * ny, true);
  d.ml_Ks = mr.in(m->ml_Ks, M, 1);
  d.ml_h = mr.in(m->ml_h, M, 1);
  d.ml_L = mr.in(m->ml_L, M, 1);
  d.ml_vpist = mr.in(m->ml_vpist, M, 1);
  d.ml_surflux = mr.in(m->ml_surflux, M, ny);
  d.ml_rest_mask = mr.in(m->ml_rest_mask, M, ny);
  d.ml_b_rest = mr.in(m->ml_b_rest, M, ny);
  d.Psi_tw = mr.out(m->Psi_tw, (size_t)M * nz, carry);
  d.Psi_iso_b = mr.out(m->Psi_iso_b, (size_t)M * nz, carry);
  d.Psi_iso_n = mr.out(m->Psi_iso_n, (size_t)M * nz, carry);
  d.psib = mr.out(m->psib, (size_t)M * nb, carry);
  d.bgrid = mr.out(m->bgrid, (size_t)M * nb, carry);
  d.Psi_so = mr.out(m->Psi_so, (size_t)M * nz, carry);
  d.Psi_Ek = mr.out(m->Psi_Ek, (size_t)M * nz, carry);
  d.Psi_GM = mr.out(m->Psi_GM, (size_t)M * nz, carry);
  d.ml_Psi_s = mr.out(m->ml_Psi_s, (size_t)M * ny, carry);
  d.status = mr.out(m->status, (size_t)M, true);
  d.scratch = nullptr;
  d.scratch_bytes = pmoc_model_scratch_bytes(m);
  if (d.scratch_bytes) d.scratch = mr.alloc((size_t)d.scratch_bytes);
  int rc = PMOC_OK;
  if (mr.err == cudaSuccess) {
    if (it0 == 0 && !(f & PMOC_ORDER_JN)) rc = pmoc_model_diagnose(&d, mr.s);
    if (rc == PMOC_OK) rc = pmoc_model_run(&d, it0, nsteps, mr.s);
    if (rc == PMOC_OK) mr.copy_back();
  }
  mr.release();
  cudaError_t e = cudaStreamSynchronize(mr.s);
  cudaStreamDestroy(mr.s);
  if (rc != PMOC_OK) return rc;
  PM_CUDA_OK(mr.err);
  PM_CUDA_OK(e);
  return PMOC_OK;
#endif
}

// ------------------------------------------------------------------------------------------
#ifndef PMOC_EMU
namespace {
// 8 independent DFMA chains per thread, nothing else: what the FP64 pipe sustains.
__global__ void k_fp64_peak(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
    a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
  }
  const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 12345.678) out[0] = s;  // keep the chains alive
}
}  // namespace
#endif

extern "C" int pmoc_fp64_peak(double* tflops, double* sm_mhz_est, void* stream) {
#ifdef PMOC_EMU
  (void)tflops; (void)sm_mhz_est; (void)stream;
  return fail(PMOC_ENODEVICE, "the FP64 probe needs a GPU");
#else
  int dev = 0, sms = 0;
  PM_CUDA_OK(cudaGetDevice(&dev));
  PM_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* out = nullptr;
  PM_CUDA_OK(cudaMalloc(&out, 8));
  cudaStream_t s = (cudaStream_t)stream;
  cudaEvent_t e0, e1;
  PM_CUDA_OK(cudaEventCreate(&e0));
  PM_CUDA_OK(cudaEventCreate(&e1));
  const int block = 512, grid = sms * 4, iters = 1 << 16;
  k_fp64_peak<<<grid, block, 0, s>>>(out, 2048, 1.0);  // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    PM_CUDA_OK(cudaEventRecord(e0, s));
    k_fp64_peak<<<grid, block, 0, s>>>(out, iters, 1.0);
    PM_CUDA_OK(cudaEventRecord(e1, s));
    PM_CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    PM_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 8.0 * (double)iters * (double)block * (double)grid;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  PM_CUDA_OK(cudaGetLastError());
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (tflops) *tflops = best;
  // 64 DFMA lanes per SM and clock
  if (sm_mhz_est) *sm_mhz_est = best * 1e12 / (2.0 * 64.0 * sms) / 1e6;
  return PMOC_OK;
#endif
}
