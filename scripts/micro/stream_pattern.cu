// Microbenchmark: what does HBM deliver for the ApplyState traffic pattern (1+4+4+4 B read, 1+2 B written per
// node) with trivial compute? Variants: classic one-shot grids vs persistent contiguous chunks.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); exit(1);} }while(0)

__device__ __forceinline__ void work(uint32_t h, uint4 f, uint4 r, uint4 d, uint32_t& n4, uint2& a4) {
  n4 = h ^ f.x ^ r.y ^ d.z; a4.x = f.y ^ r.z ^ d.w; a4.y = f.w ^ r.x ^ d.y ^ f.z ^ r.w ^ d.x;
}
// C: one group of 4 nodes per thread, classic grid
template <int G>
__global__ void classic(long long n4, const uint32_t* h, const uint4* f, const uint4* r, const uint4* d, uint32_t* nx, uint2* ac) {
  long long q0 = ((long long)blockIdx.x * blockDim.x) * G + threadIdx.x;
  uint32_t hh[G]; uint4 ff[G], rr[G], dd[G];
#pragma unroll
  for (int g = 0; g < G; g++) { long long q = q0 + (long long)g * blockDim.x; if (q < n4) { hh[g] = __ldg(h + q); ff[g] = __ldcs(f + q); rr[g] = __ldcs(r + q); dd[g] = __ldcs(d + q); } }
#pragma unroll
  for (int g = 0; g < G; g++) { long long q = q0 + (long long)g * blockDim.x; if (q < n4) { uint32_t n; uint2 a; work(hh[g], ff[g], rr[g], dd[g], n, a); __stcs(nx + q, n); __stcs(ac + q, a); } }
}
// P: persistent contiguous chunks, U groups per thread per iteration, optional double buffering
template <int U, bool DB>
__global__ void persistent(long long n4, const uint32_t* h, const uint4* f, const uint4* r, const uint4* d, uint32_t* nx, uint2* ac) {
  long long per = (n4 + gridDim.x - 1) / gridDim.x; per = (per + 31) & ~31LL;
  long long b0 = per * blockIdx.x, b1 = b0 + per < n4 ? b0 + per : n4;
  const int T = blockDim.x;
  uint32_t hh[2][U]; uint4 ff[2][U], rr[2][U], dd[2][U];
  auto load = [&](int s, long long base) {
#pragma unroll
    for (int u = 0; u < U; u++) { long long q = base + (long long)u * T + threadIdx.x; if (q < b1) { hh[s][u] = __ldg(h + q); ff[s][u] = __ldcs(f + q); rr[s][u] = __ldcs(r + q); dd[s][u] = __ldcs(d + q); } } };
  auto comp = [&](int s, long long base) {
#pragma unroll
    for (int u = 0; u < U; u++) { long long q = base + (long long)u * T + threadIdx.x; if (q < b1) { uint32_t n; uint2 a; work(hh[s][u], ff[s][u], rr[s][u], dd[s][u], n, a); __stcs(nx + q, n); __stcs(ac + q, a); } } };
  const long long step = (long long)U * T;
  if (DB) {
    load(0, b0);
    for (long long base = b0; base < b1; base += 2 * step) { load(1, base + step); comp(0, base); load(0, base + 2 * step); comp(1, base + step); }
  } else {
    for (long long base = b0; base < b1; base += step) { load(0, base); comp(0, base); }
  }
}
int main() {
  const long long n = 10000000, n4 = n / 4; const int SETS = 8;
  std::vector<void*> H(SETS), F(SETS), R(SETS), D(SETS), N(SETS), A(SETS);
  for (int s = 0; s < SETS; s++) { CK(cudaMalloc(&H[s], n)); CK(cudaMalloc(&F[s], n * 4)); CK(cudaMalloc(&R[s], n * 4)); CK(cudaMalloc(&D[s], n * 4)); CK(cudaMalloc(&N[s], n)); CK(cudaMalloc(&A[s], n * 2));
    CK(cudaMemset(H[s], 1, n)); CK(cudaMemset(F[s], 2, n * 4)); CK(cudaMemset(R[s], 3, n * 4)); CK(cudaMemset(D[s], 4, n * 4)); }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  auto run = [&](const char* name, auto launch) {
    std::vector<float> t;
    for (int it = 0; it < 30; it++) { int s = it % SETS; cudaEventRecord(e0); launch(s); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (it >= 5) t.push_back(ms); }
    std::sort(t.begin(), t.end()); float med = t[t.size() / 2];
    printf("%-44s %8.2f us  %7.1f GB/s\n", name, med * 1e3, 16.0 * n / (med * 1e-3) / 1e9);
  };
#define ARGS(s) n4, (const uint32_t*)H[s], (const uint4*)F[s], (const uint4*)R[s], (const uint4*)D[s], (uint32_t*)N[s], (uint2*)A[s]
  run("classic G=1 256thr", [&](int s) { classic<1><<<(unsigned)((n4 + 255) / 256), 256>>>(ARGS(s)); });
  run("classic G=2 256thr", [&](int s) { classic<2><<<(unsigned)((n4 + 511) / 512), 256>>>(ARGS(s)); });
  run("classic G=4 256thr", [&](int s) { classic<4><<<(unsigned)((n4 + 1023) / 1024), 256>>>(ARGS(s)); });
  run("classic G=1 1024thr", [&](int s) { classic<1><<<(unsigned)((n4 + 1023) / 1024), 1024>>>(ARGS(s)); });
  run("persistent U=2 noDB 296x256", [&](int s) { persistent<2, false><<<2 * sms, 256>>>(ARGS(s)); });
  run("persistent U=2 DB   296x256", [&](int s) { persistent<2, true><<<2 * sms, 256>>>(ARGS(s)); });
  run("persistent U=4 noDB 296x256", [&](int s) { persistent<4, false><<<2 * sms, 256>>>(ARGS(s)); });
  run("persistent U=2 DB   592x256", [&](int s) { persistent<2, true><<<4 * sms, 256>>>(ARGS(s)); });
  run("persistent U=2 DB   1184x256", [&](int s) { persistent<2, true><<<8 * sms, 256>>>(ARGS(s)); });
  run("persistent U=1 noDB 1184x256", [&](int s) { persistent<1, false><<<8 * sms, 256>>>(ARGS(s)); });
  run("persistent U=2 DB   296x512", [&](int s) { persistent<2, true><<<2 * sms, 512>>>(ARGS(s)); });
  run("persistent U=2 DB   148x1024", [&](int s) { persistent<2, true><<<sms, 1024>>>(ARGS(s)); });
  run("persistent U=4 DB   148x1024", [&](int s) { persistent<4, true><<<sms, 1024>>>(ARGS(s)); });
  return 0;
}
