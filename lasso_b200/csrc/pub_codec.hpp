// lasso_b200 — the wire format of a tagged publication (common.cuh PubDst, prover.cu Ctx::pub_wait_raw): a value
// x < 2^255 (8 x u32 little-endian: an Fr residue or a canonical Fq coordinate) as FIVE 64-bit words, word k = bits
// [51k, 51k + 51) of x in its low 51 bits and the 13-bit tag of its message in its high bits.  Every word identifies
// its message on its own, so the receiver needs nothing stronger than the single-copy atomicity of an aligned 64-bit
// store.  Shared by the device-side encoder and the host-side decoder (and tested on the CPU, tests/test_product_host.py).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define LB_PUB_HD __host__ __device__ __forceinline__
#else
#define LB_PUB_HD inline
#endif

namespace lb {

static constexpr unsigned long long kPubValueMask = (1ull << 51) - 1;

LB_PUB_HD void pub_encode(const uint32_t x[8], uint32_t tag, unsigned long long w[5]) {
  const unsigned long long q0 = x[0] | ((unsigned long long)x[1] << 32), q1 = x[2] | ((unsigned long long)x[3] << 32),
                           q2 = x[4] | ((unsigned long long)x[5] << 32), q3 = x[6] | ((unsigned long long)x[7] << 32);
  const unsigned long long T = (unsigned long long)tag << 51;
  w[0] = (q0 & kPubValueMask) | T;
  w[1] = (((q0 >> 51) | (q1 << 13)) & kPubValueMask) | T;
  w[2] = (((q1 >> 38) | (q2 << 26)) & kPubValueMask) | T;
  w[3] = (((q2 >> 25) | (q3 << 39)) & kPubValueMask) | T;
  w[4] = (q3 >> 12) | T;
}
// the tag a word carries (0 = empty slot)
LB_PUB_HD uint32_t pub_tag_of(unsigned long long word) { return (uint32_t)(word >> 51); }
// w: the five words with their tags already stripped (& kPubValueMask)
LB_PUB_HD void pub_decode(const unsigned long long w[5], uint32_t x[8]) {
  const unsigned long long q0 = w[0] | (w[1] << 51), q1 = (w[1] >> 13) | (w[2] << 38), q2 = (w[2] >> 26) | (w[3] << 25),
                           q3 = (w[3] >> 39) | (w[4] << 12);
  x[0] = (uint32_t)q0;
  x[1] = (uint32_t)(q0 >> 32);
  x[2] = (uint32_t)q1;
  x[3] = (uint32_t)(q1 >> 32);
  x[4] = (uint32_t)q2;
  x[5] = (uint32_t)(q2 >> 32);
  x[6] = (uint32_t)q3;
  x[7] = (uint32_t)(q3 >> 32);
}

}  // namespace lb
