// Host-buffer entry point (pmoc_model_run_host) and the FP64 roofline probe.
#include "pmoc_common.cuh"

#include <vector>

static thread_local unsigned long long g_h2d_bytes = 0, g_d2h_bytes = 0;

extern "C" void pmoc_host_last_bytes(uint64_t* h2d, uint64_t* d2h) {
  if (h2d) *h2d = g_h2d_bytes;
  if (d2h) *d2h = g_d2h_bytes;
}

#ifndef PMOC_EMU
namespace {

// Mirrors the host arrays of a pmoc_model on the device for the duration of one call, and pipelines the
// call over blocks of members: block c's host->device copies, kernel and device->host copies are queued on
// stream c % kStreams, so the copies of one block overlap the kernel of another (members are independent).
constexpr int kStreams = 3;
#ifndef PMOC_HOST_BLOCK_MEMBERS
#define PMOC_HOST_BLOCK_MEMBERS 4096  // 16 blocks for the 65,536-member bench: fill/drain of the pipeline ~6 % of the copies
#endif

struct Mirror {
  struct Field {
    char* host;        // host base
    char* dev;         // device base
    size_t slot;       // byte offset of the pointer inside pmoc_model
    size_t per_member; // bytes per member; 0 = shared by all members
    size_t bytes;      // total bytes
    bool in, out;      // copied in / copied back
  };
  cudaStream_t s[kStreams] = {};
  cudaEvent_t shared_ready = nullptr;
  std::vector<Field> fields;
  std::vector<void*> allocs;
  cudaError_t err = cudaSuccess;
  const pmoc_model* h;
  pmoc_model d;

  void* alloc(size_t bytes) {
    void* p = nullptr;
    if (err == cudaSuccess) err = cudaMallocAsync(&p, bytes ? bytes : 8, s[0]);
    if (p) allocs.push_back(p);
    return p;
  }
  // register the pointer stored at `slot` of the model (hp: its host value)
  template <class P>
  void add(P* slot, const void* hp, size_t per_member, size_t bytes, bool in, bool out) {
    if (!hp) return;
    char* dp = static_cast<char*>(alloc(bytes));
    fields.push_back({(char*)hp, dp, (size_t)((char*)slot - (char*)&d), per_member, bytes, in, out});
    *slot = reinterpret_cast<P>(dp);
  }
  void vec(pmoc_vec* dv, long long M, size_t len) {  // input vector: per member when mstride != 0
    if (!dv->ptr) return;
    const size_t stride = (size_t)dv->mstride * sizeof(double);
    const size_t bytes = stride ? (size_t)(M - 1) * stride + len * sizeof(double) : len * sizeof(double);
    add(&dv->ptr, dv->ptr, stride, bytes, true, false);
  }
  template <class T>
  void shared(const T** dp, size_t n) { add(dp, *dp, 0, n * sizeof(T), true, false); }
  template <class T>
  void state(T** dp, long long M, size_t len, bool upload, bool download = true) {  // per-member in/out array
    add(dp, *dp, len * sizeof(T), (size_t)M * len * sizeof(T), upload, download);
  }
  void column(pmoc_column* c, long long M, int nz) {
    state(&c->b, M, nz, true);
    vec(&c->kappa, M, (size_t)c->nvar * nz);
    vec(&c->dAk, M, (size_t)c->nvar * nz);
    vec(&c->Area, M, nz);
    vec(&c->bs, M, 1);
    vec(&c->N2min, M, 1);
    vec(&c->bzbot, M, 1);
    state(&c->bbot, M, 1, true);
    state(&c->var, M, 1, true);
  }
  void upload_shared() {
    for (auto& f : fields)
      if (!f.per_member && f.in && err == cudaSuccess) {
        err = cudaMemcpyAsync(f.dev, f.host, f.bytes, cudaMemcpyHostToDevice, s[0]);
        g_h2d_bytes += f.bytes;
      }
    if (err == cudaSuccess) err = cudaEventRecord(shared_ready, s[0]);
  }
  // members [m0, m0+n) in, on stream st; the last member of an input vector may be shorter than its stride
  void upload_block(long long m0, long long n, long long M, cudaStream_t st) {
    for (auto& f : fields) {
      if (!f.per_member || err != cudaSuccess) continue;
      const size_t off = (size_t)m0 * f.per_member;
      size_t len = (size_t)n * f.per_member;
      if (off + len > f.bytes) len = f.bytes - off;
      err = f.in ? cudaMemcpyAsync(f.dev + off, f.host + off, len, cudaMemcpyHostToDevice, st)
                 : cudaMemsetAsync(f.dev + off, 0, len, st);
      if (f.in) g_h2d_bytes += len;
    }
  }
  void download_block(long long m0, long long n, cudaStream_t st) {
    for (auto& f : fields) {
      if (!f.per_member || !f.out || err != cudaSuccess) continue;
      const size_t off = (size_t)m0 * f.per_member;
      err = cudaMemcpyAsync(f.host + off, f.dev + off, (size_t)n * f.per_member, cudaMemcpyDeviceToHost, st);
      g_d2h_bytes += (size_t)n * f.per_member;
    }
  }
  // the device model restricted to members [m0, m0+n)
  pmoc_model block_model(long long m0, long long n) const {
    pmoc_model b = d;
    b.M = n;
    for (auto& f : fields)
      if (f.per_member) *reinterpret_cast<char**>((char*)&b + f.slot) = f.dev + (size_t)m0 * f.per_member;
    return b;
  }
  void release() {
    for (void* p : allocs) cudaFreeAsync(p, s[0]);
    allocs.clear();
  }
};

}  // namespace
#endif

extern "C" int pmoc_model_run_host(const pmoc_model* m, int64_t it0, int64_t nsteps) {
  if (!m) return fail(PMOC_EINVAL, "model is NULL");
  g_h2d_bytes = g_d2h_bytes = 0;
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
  for (int i = 0; i < kStreams; ++i) PM_CUDA_OK(cudaStreamCreateWithFlags(&mr.s[i], cudaStreamNonBlocking));
  PM_CUDA_OK(cudaEventCreateWithFlags(&mr.shared_ready, cudaEventDisableTiming));
  mr.h = m;
  mr.d = *m;
  pmoc_model& d = mr.d;
  // Diagnostics: the kernel rewrites all of them at the last iteration with ii % K == 0 of the launch (and the
  // pre-loop diagnosis of an order-'post' model at it0 == 0 does); without such an iteration they are left
  // alone on both sides.  Only the streamfunctions the step loop carries from an earlier launch are inputs.
  const long long K = m->K > 0 ? m->K : 1;
  const bool jn = (f & PMOC_ORDER_JN) != 0;
  const bool rewritten = (!jn && it0 == 0) || (nsteps > 0 && ((it0 + nsteps - 1) / K) * K >= it0);
  const bool carry = jn ? (it0 % K != 0) : (it0 > 0);
  mr.shared(&d.z, nz);
  mr.shared(&d.y, ny);
  mr.column(&d.basin, M, nz);
  if (f & PMOC_HAS_NORTH) mr.column(&d.north, M, nz);
  if (f & PMOC_HAS_PAC) {
    mr.column(&d.pac, M, nz);
    mr.vec(&d.zoc_f, M, 1);
    mr.vec(&d.so2_L, M, 1);
    mr.state(&d.Psi_zoc, M, nz, false, rewritten);
    mr.state(&d.Psi_zon_a, M, nz, carry, rewritten);
    mr.state(&d.Psi_zon_p, M, nz, carry, rewritten);
    mr.state(&d.psib2, M, nb, false, rewritten);
    mr.state(&d.bgrid2, M, nb, false, rewritten);
    mr.state(&d.Psi_so2, M, nz, carry, rewritten);
    mr.state(&d.Psi_Ek2, M, nz, false, rewritten);
    mr.state(&d.Psi_GM2, M, nz, false, rewritten);
  }
  mr.vec(&d.tw_f, M, 1);
  mr.vec(&d.tw_b2, M, nz);
  mr.vec(&d.so_bs, M, ny);
  mr.vec(&d.so_tau, M, m->so_tau_on_y ? ny : 1);
  mr.vec(&d.so_f, M, 1);
  mr.vec(&d.so_rho, M, 1);
  mr.vec(&d.so_L, M, 1);
  mr.vec(&d.so_KGM, M, 1);
  mr.vec(&d.so_smax, M, 1);
  mr.vec(&d.so_c, M, 1);
  mr.shared(&d.so_sill_taper, nz);
  mr.shared(&d.so_ek_taper, nz);
  mr.shared(&d.so_top_taper, nz);
  mr.shared(&d.so_bot_taper, nz);
  mr.state(&d.ml_bs, M, ny, true);
  mr.vec(&d.ml_Ks, M, 1);
  mr.vec(&d.ml_h, M, 1);
  mr.vec(&d.ml_L, M, 1);
  mr.vec(&d.ml_vpist, M, 1);
  mr.vec(&d.ml_surflux, M, ny);
  mr.vec(&d.ml_rest_mask, M, ny);
  mr.vec(&d.ml_b_rest, M, ny);
  mr.state(&d.Psi_tw, M, nz, carry && !(f & PMOC_ISO), rewritten);
  mr.state(&d.Psi_iso_b, M, nz, carry, rewritten);
  mr.state(&d.Psi_iso_n, M, nz, carry, rewritten);
  mr.state(&d.psib, M, nb, false, rewritten);
  mr.state(&d.bgrid, M, nb, false, rewritten);
  mr.state(&d.Psi_so, M, nz, carry, rewritten);
  mr.state(&d.Psi_Ek, M, nz, false, rewritten);
  mr.state(&d.Psi_GM, M, nz, false, rewritten);
  mr.state(&d.ml_Psi_s, M, ny, false, nsteps > 0);
  mr.state(&d.status, M, 1, true);
  // blocks of members: enough of them to overlap copies with kernels, each large enough to fill the GPU
  long long nblk = M / PMOC_HOST_BLOCK_MEMBERS;
  if (nblk < 1) nblk = 1;
  if (nblk > 32) nblk = 32;
  const long long per = (M + nblk - 1) / nblk;
  void* scratch[kStreams] = {};
  d.scratch = nullptr;
  d.scratch_bytes = 0;
  {  // block-per-member kernels (nz > 256): one scratch buffer per stream, sized for a block
    pmoc_model one = *m;
    one.M = per;
    const uint64_t need = pmoc_model_scratch_bytes(&one);
    if (need) {
      for (int i = 0; i < kStreams && i < nblk; ++i) scratch[i] = mr.alloc((size_t)need);
      d.scratch_bytes = need;
    }
  }
  int rc = PMOC_OK;
  if (mr.err == cudaSuccess) {
    mr.upload_shared();
    for (int i = 1; i < kStreams && mr.err == cudaSuccess; ++i) mr.err = cudaStreamWaitEvent(mr.s[i], mr.shared_ready, 0);
    for (long long c = 0; c < nblk && rc == PMOC_OK && mr.err == cudaSuccess; ++c) {
      const long long m0 = c * per, n = (m0 + per <= M ? per : M - m0);
      if (n <= 0) break;
      cudaStream_t st = mr.s[c % kStreams];
      mr.upload_block(m0, n, M, st);
      pmoc_model blk = mr.block_model(m0, n);
      blk.scratch = scratch[c % kStreams];
      if (it0 == 0 && !(f & PMOC_ORDER_JN)) rc = pmoc_model_diagnose(&blk, st);
      if (rc == PMOC_OK) rc = pmoc_model_run(&blk, it0, nsteps, st);
      if (rc == PMOC_OK) mr.download_block(m0, n, st);
    }
  }
  cudaError_t e = cudaSuccess;
  for (int i = kStreams - 1; i >= 0; --i) {  // stream 0 last: it owns the allocations
    if (i == 0) mr.release();
    const cudaError_t ei = cudaStreamSynchronize(mr.s[i]);
    if (e == cudaSuccess) e = ei;
  }
  for (int i = 0; i < kStreams; ++i) cudaStreamDestroy(mr.s[i]);
  cudaEventDestroy(mr.shared_ready);
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
