// Microbenchmark: GPU-side cost of a cooperative launch vs a regular launch (event-to-event, back-to-back stream of launches)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <algorithm>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;} }while(0)
__global__ void __launch_bounds__(256, 2) k_empty(unsigned* p) { if (p == (unsigned*)1) *p = 0; }
__global__ void __launch_bounds__(256, 2) k_smem(unsigned* p) { __shared__ unsigned s[10000]; s[threadIdx.x] = threadIdx.x; __syncthreads(); if (p == (unsigned*)1) *p = s[5]; }
__global__ void __launch_bounds__(256, 2) k_barrier(unsigned* ctr, unsigned n) {
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); atomicAdd(ctr, 1u); unsigned v; do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < n); }
  __syncthreads();
  if (threadIdx.x == 0) { unsigned prev = atomicAdd(ctr + 1, 1u); if (prev == n - 1) { ctr[0] = 0; ctr[1] = 0; } }
}
int main() {
  unsigned* ctr; CK(cudaMalloc(&ctr, 64)); CK(cudaMemset(ctr, 0, 64));
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  const int R = 200;
  auto timeit = [&](const char* name, auto launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 20; i++) launch();
    cudaStreamSynchronize(st);
    cudaEventRecord(e0, st); for (int i = 0; i < R; i++) launch(); cudaEventRecord(e1, st); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); printf("%-46s %7.2f us per launch (stream of %d)\n", name, ms * 1e3 / R, R);
    return 0;
  };
  unsigned n = 296; void* args[] = {&ctr, &n}; void* a1[] = {&ctr};
  timeit("regular empty 296x256", [&] { k_empty<<<296, 256, 0, st>>>(ctr); });
  timeit("cooperative empty 296x256", [&] { cudaLaunchCooperativeKernel((const void*)k_empty, dim3(296), dim3(256), a1, 0, st); });
  timeit("regular 40KB-smem 296x256", [&] { k_smem<<<296, 256, 0, st>>>(ctr); });
  timeit("cooperative 40KB-smem 296x256", [&] { cudaLaunchCooperativeKernel((const void*)k_smem, dim3(296), dim3(256), a1, 0, st); });
  timeit("cooperative grid-barrier 296x256", [&] { cudaLaunchCooperativeKernel((const void*)k_barrier, dim3(296), dim3(256), args, 0, st); });
  timeit("regular grid-barrier 296x256 (unsafe)", [&] { k_barrier<<<296, 256, 0, st>>>(ctr, n); });
  timeit("2x regular empty back-to-back", [&] { k_empty<<<296, 256, 0, st>>>(ctr); k_empty<<<296, 256, 0, st>>>(ctr); });
  return 0;
}
