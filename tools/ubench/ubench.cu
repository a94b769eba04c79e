// Microbenchmarks that size the latency-bound parts of the prover (not part of the product):
// dependent-chain latency and multi-warp throughput of IMAD.WIDE, fq_mul, fr_mul, pt_add on one SM.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../lasso_b200/csrc/ed25519.cuh"
using namespace lb;

__global__ void k_imad_dep(uint64_t* out, int iters, uint32_t a, uint32_t b) {
  uint64_t x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x) : "r"(a), "r"(b));
  }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[4096] = (uint64_t)(t1 - t0);
}
template <int ILP>
__global__ void k_imad_ilp(uint64_t* out, int iters, uint32_t a, uint32_t b) {
  uint64_t x[ILP];
  for (int j = 0; j < ILP; j++) x[j] = threadIdx.x + j;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++)
#pragma unroll
      for (int j = 0; j < ILP; j++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[j]) : "r"(a), "r"(b));
  }
  long long t1 = clock64();
  uint64_t s = 0;
  for (int j = 0; j < ILP; j++) s += x[j];
  out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[4096] = (uint64_t)(t1 - t0);
}
template <int ILP>
__global__ void k_fq_mul(uint32_t* out, int iters) {
  fq_t x[ILP], y;
  for (int j = 0; j < ILP; j++)
    for (int l = 0; l < 8; l++) x[j].v[l] = threadIdx.x * 77u + l * 12345u + j;
  for (int l = 0; l < 8; l++) y.v[l] = 0x9e3779b9u * (l + 1) + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < ILP; j++) x[j] = fq_mul(x[j], y);
  }
  long long t1 = clock64();
  uint32_t s = 0;
  for (int j = 0; j < ILP; j++)
    for (int l = 0; l < 8; l++) s ^= x[j].v[l];
  out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) ((uint64_t*)out)[4096] = (uint64_t)(t1 - t0);
}
template <int ILP>
__global__ void k_fr_mul(uint32_t* out, int iters) {
  fr_t x[ILP], y;
  for (int j = 0; j < ILP; j++)
    for (int l = 0; l < 8; l++) x[j].v[l] = (threadIdx.x * 77u + l * 12345u + j) & 0x0fffffffu;
  for (int l = 0; l < 8; l++) y.v[l] = (0x9e3779b9u * (l + 1) + threadIdx.x) & 0x0fffffffu;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < ILP; j++) x[j] = fr_mul(x[j], y);
  }
  long long t1 = clock64();
  uint32_t s = 0;
  for (int j = 0; j < ILP; j++)
    for (int l = 0; l < 8; l++) s ^= x[j].v[l];
  out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) ((uint64_t*)out)[4096] = (uint64_t)(t1 - t0);
}
__global__ void k_pt_add(uint32_t* out, int iters) {
  pt_ext p = pt_identity(), q = pt_identity();
  q.X.v[0] = 5 + threadIdx.x;
  q.T.v[0] = 7;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) p = pt_add(p, q);
  long long t1 = clock64();
  uint32_t s = 0;
  for (int l = 0; l < 8; l++) s ^= p.X.v[l] ^ p.Y.v[l] ^ p.Z.v[l] ^ p.T.v[l];
  out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) ((uint64_t*)out)[4096] = (uint64_t)(t1 - t0);
}

template <typename F>
static void run(const char* name, F launch, int iters, double per_iter_ops, void* d) {
  launch();
  cudaDeviceSynchronize();
  launch();
  cudaError_t e = cudaDeviceSynchronize();
  uint64_t cyc = 0;
  cudaMemcpy(&cyc, (uint64_t*)d + 4096, 8, cudaMemcpyDeviceToHost);
  printf("%-44s %10.1f cycles/op (%s)\n", name, (double)cyc / (iters * per_iter_ops), cudaGetErrorString(e));
}
int main() {
  void* d;
  cudaMalloc(&d, 1 << 20);
  const int it = 2000;
  for (int warps : {1, 4, 8, 16, 32}) {
    int th = warps * 32;
    printf("--- %d warp(s) on one SM (%d per scheduler)\n", warps, (warps + 3) / 4);
    run("mad.wide.u32 dependent chain", [&] { k_imad_dep<<<1, th>>>((uint64_t*)d, it, 3, 5); }, it, 16, d);
    run("mad.wide.u32 4 independent chains", [&] { k_imad_ilp<4><<<1, th>>>((uint64_t*)d, it, 3, 5); }, it, 64, d);
    run("mad.wide.u32 8 independent chains", [&] { k_imad_ilp<8><<<1, th>>>((uint64_t*)d, it, 3, 5); }, it, 128, d);
    run("fq_mul dependent", [&] { k_fq_mul<1><<<1, th>>>((uint32_t*)d, it); }, it, 1, d);
    run("fq_mul 2 independent / thread", [&] { k_fq_mul<2><<<1, th>>>((uint32_t*)d, it); }, it, 2, d);
    run("fq_mul 4 independent / thread", [&] { k_fq_mul<4><<<1, th>>>((uint32_t*)d, it); }, it, 4, d);
    run("fr_mul dependent", [&] { k_fr_mul<1><<<1, th>>>((uint32_t*)d, it); }, it, 1, d);
    run("fr_mul 2 independent / thread", [&] { k_fr_mul<2><<<1, th>>>((uint32_t*)d, it); }, it, 2, d);
    run("fr_mul 4 independent / thread", [&] { k_fr_mul<4><<<1, th>>>((uint32_t*)d, it); }, it, 4, d);
    run("pt_add dependent", [&] { k_pt_add<<<1, th>>>((uint32_t*)d, it / 4); }, it / 4, 1, d);
  }
  return 0;
}
