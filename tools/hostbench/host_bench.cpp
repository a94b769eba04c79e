// Host-side microbenchmarks of what sits between two kernel launches of a proof: the Keccak-f[1600] permutation behind
// the Merlin transcript (portable build against the run-time-dispatched x86-64-v3 build), the field inversions of an
// opening round (binary GCD against the exponentiation), and the two together as one opening round's host work.
//   g++ -O3 -std=c++17 -I lasso_b200/csrc tools/hostbench/host_bench.cpp -o /tmp/host_bench && /tmp/host_bench
#include <chrono>
#include <cstdio>
#include <vector>

#include "host_fq64.hpp"
#include "host_transcript.hpp"
using namespace lb;
template <class F>
static double us_per(int iters, F f) {
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < iters; i++) f(i);
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / iters;
}
int main() {
  uint64_t st[25] = {1, 2, 3};
  printf("keccak-f[1600] portable   : %.3f us\n", us_per(200000, [&](int) { KeccakF1600::permute_portable(st); }));
  printf("keccak-f[1600] dispatched : %.3f us\n", us_per(200000, [&](int) { KeccakF1600::permute(st); }));
  Transcript t("bench");
  std::vector<uint8_t> a(2048 * 32, 7);
  printf("append_scalars_bytes(\"a\", 2048 scalars): %.1f us\n", us_per(200, [&](int) { t.append_scalars_bytes("a", a.data(), 2048); }));
  uint32_t p[24], q[24];
  for (int i = 0; i < 24; i++) {
    p[i] = 0x9e3779b9u * (i + 1) + (uint32_t)st[0];
    q[i] = 0x85ebca6bu * (i + 3);
  }
  for (int k : {7, 15, 23}) p[k] &= 0x7fffffffu, q[k] &= 0x7fffffffu;
  uint8_t oa[32], ob[32];
  printf("compress (L, R) pair, binary-GCD inversion : %.3f us\n", us_per(50000, [&](int) {
           h64::compress_xyz_pair(p, q, oa, ob);
           p[0] ^= oa[0];
         }));
  h64::fe z = h64::from_limbs32(p + 16);
  printf("Fq inversion by exponentiation             : %.3f us\n", us_per(20000, [&](int) { z = h64::inv_fermat(z); }));
  fr_t x = fr_from_u64(12345);
  printf("Fr inversion, binary GCD                   : %.3f us\n", us_per(50000, [&](int) { x = fr_add(fr_inv(x), fr_one()); }));
  printf("Fr inversion by exponentiation             : %.3f us\n", us_per(20000, [&](int) { x = fr_add(frh::inv_fermat(x), fr_one()); }));
  uint8_t LR[64] = {0};
  fr_t u = fr_one();
  printf("one opening round on the host (compress pair, append L, R, challenge u, u^-1): %.3f us\n", us_per(50000, [&](int) {
           h64::compress_xyz_pair(p, q, LR, LR + 32);
           t.append_point_compressed("L", LR);
           t.append_point_compressed("R", LR + 32);
           u = t.challenge_scalar("u");
           u = fr_inv(u);
           p[0] ^= u.v[0];
         }));
  printf("%u %llu\n", x.v[0] ^ u.v[1], (unsigned long long)(st[0] ^ z.v[0]));
  return 0;
}
