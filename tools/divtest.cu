#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "../pymoc_b200/csrc/pmoc_rt.cuh"
__global__ void k(const double* a, const double* b, long long n, unsigned long long* bad) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) { double q = rt::div_normal(a[i], b[i]); double w = a[i] / b[i]; if (__double_as_longlong(q) != __double_as_longlong(w)) atomicAdd(bad, 1ull); }
}
int main() {
  const long long n = 1 << 26;
  double *a, *b; unsigned long long* bad;
  cudaMallocManaged(&a, n * 8); cudaMallocManaged(&b, n * 8); cudaMallocManaged(&bad, 8); *bad = 0;
  srand48(1);
  for (long long i = 0; i < n; ++i) {
    int mode = i & 3;
    if (mode == 0) { a[i] = -4000.0 * drand48(); b[i] = 0.1 + 2e6 * drand48(); }
    else if (mode == 1) { a[i] = (drand48() - 0.5) * 60; b[i] = 6e-5 * (1 + 1e-9 * drand48()); }
    else if (mode == 2) { a[i] = ldexp(drand48() - 0.5, (int)(600 * drand48()) - 300); b[i] = ldexp(0.5 + drand48(), (int)(600 * drand48()) - 300); }
    else { a[i] = (i & 4) ? 0.0 : -0.0; b[i] = drand48() + 1e-3; }
  }
  k<<<(n + 255) / 256, 256>>>(a, b, n, bad);
  cudaDeviceSynchronize();
  printf("mismatches vs IEEE divide: %llu of %lld\n", *bad, n);
  return 0;
}
