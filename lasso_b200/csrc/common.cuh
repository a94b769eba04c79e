// lasso_b200 — shared device helpers: 256-bit global loads/stores of field elements,
// warp-shuffle + shared-memory reductions of Fr partial sums, error checking.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <stdexcept>
#include <string>

#include "ed25519.cuh"

namespace lb {

#define LB_CUDA_CHECK(x)                                                                              \
  do {                                                                                                \
    cudaError_t e_ = (x);                                                                             \
    if (e_ != cudaSuccess)                                                                            \
      throw std::runtime_error(std::string("CUDA error ") + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + \
                               std::to_string(__LINE__));                                             \
  } while (0)

static constexpr int kNumSMs = 148;  // B200

#if defined(__CUDACC__)
// One 32-byte element per thread per instruction: LDG.E.ENL2.256 / STG.E.ENL2.256 on sm_100a,
// so a warp moves 1 KiB fully coalesced.
__device__ __forceinline__ fr_t ld_fr(const fr_t* p) {
  fr_t r;
  asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
                 "=r"(r.v[7])
               : "l"(p));
  return r;
}
// streaming variant for data read exactly once per kernel
__device__ __forceinline__ fr_t ld_fr_stream(const fr_t* p) {
  fr_t r;
  asm volatile("ld.global.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
                 "=r"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_fr(fr_t* p, const fr_t& r) {
  asm volatile("st.global.v8.u32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]),
               "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7]), "l"(p)
               : "memory");
}
__device__ __forceinline__ fq_t ld_fq(const fq_t* p) {
  fq_t r;
  asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
                 "=r"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_fq(fq_t* p, const fq_t& r) {
  asm volatile("st.global.v8.u32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]),
               "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7]), "l"(p)
               : "memory");
}

__device__ __forceinline__ fr_t shfl_down_fr(const fr_t& a, int delta) {
  fr_t r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_down_sync(0xffffffffu, a.v[i], delta);
  return r;
}
// sum over the warp, result valid in lane 0
__device__ __forceinline__ fr_t warp_sum_fr(fr_t a) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) a = fr_add(a, shfl_down_fr(a, d));
  return a;
}
// Sum NV values per thread over the block.  `scratch` needs NV * (blockDim.x/32) elements.
// Result valid in thread 0.
template <int NV>
__device__ __forceinline__ void block_sum_fr(fr_t (&v)[NV], fr_t* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    v[k] = warp_sum_fr(v[k]);
    if (lane == 0) scratch[k * nwarps + warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) {
      fr_t x = lane < nwarps ? scratch[k * nwarps + lane] : fr_zero();
      v[k] = warp_sum_fr(x);
    }
  }
  __syncthreads();
}
#endif

}  // namespace lb
