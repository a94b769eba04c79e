// lasso_b200 — host-side modular inversion for the two 255-bit prime fields of the path (Fr: the scalar field l,
// Fq: 2^255 - 19), used where an inversion sits on the critical path between two kernel launches: every Bulletproofs
// round needs u^-1 of the challenge it just squeezed (bullet.rs:100-101) and the affine form of the L, R points it
// absorbs (ark-serialize compresses normalised points).  Fermat's a^(p-2) is ~320 dependent 256-bit
// multiplications (~7 us here); this is the binary extended GCD in the batched form of T. Pornin, "Optimized Binary
// GCD for Modular Inversion" (2020): 31 plain binary-GCD steps are run on 64-bit approximations of (a, b) — their
// 31 low bits, which decide parities exactly, and their 33 top bits, which decide the comparisons — while the 2x2
// update matrix of those steps is collected in machine words; the matrix is then applied once to the full-width
// (a, b) and, modulo m, to the Bezout coefficients (u, v).  ~15 such rounds for 255-bit operands.  The operands are
// public (challenges, published points): no constant-time requirement, the loop stops when a = 0.
// Results are exact residues, so callers get the same bits as from the exponentiation (tests/test_product_host.py).
#pragma once
#include <cstdint>
#include <cstring>

namespace lb {
namespace modinv {

typedef unsigned __int128 u128;

struct Modulus {
  uint64_t m[4];
  uint64_t ninv31;  // -m^-1 mod 2^31
};
inline Modulus make_modulus(const uint64_t m[4]) {
  Modulus M;
  memcpy(M.m, m, 32);
  uint64_t x = m[0];  // Newton: x <- x (2 - m x) doubles the number of correct low bits (m odd: 3 to start with)
  for (int i = 0; i < 5; i++) x *= 2 - m[0] * x;
  M.ninv31 = (0 - x) & 0x7fffffffULL;
  return M;
}

inline int bitlen4(const uint64_t a[4]) {
  for (int i = 3; i >= 0; i--)
    if (a[i]) return 64 * i + 64 - __builtin_clzll(a[i]);
  return 0;
}
inline bool is_zero4(const uint64_t a[4]) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
inline uint64_t bits33_at(const uint64_t a[4], int pos) {  // bits [pos, pos + 33) of a, 0 <= pos <= 223
  const int w = pos >> 6, off = pos & 63;
  uint64_t v = a[w] >> off;
  if (off > 31 && w + 1 < 4) v |= a[w + 1] << (64 - off);
  return v & 0x1ffffffffULL;
}
// out (5 words, two's complement) = f * a, f signed with |f| <= 2^31
inline void mul_signed(uint64_t out[5], const uint64_t a[4], int64_t f) {
  const uint64_t mag = f < 0 ? (uint64_t)0 - (uint64_t)f : (uint64_t)f;
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a[i] * mag;
    out[i] = (uint64_t)c;
    c >>= 64;
  }
  out[4] = (uint64_t)c;
  if (f < 0) {
    u128 b = 1;
    for (int i = 0; i < 5; i++) {
      b += (uint64_t)~out[i];
      out[i] = (uint64_t)b;
      b >>= 64;
    }
  }
}
// r = (f a + g b) / 2^31 as a magnitude (4 words) and a sign; the division is exact by construction
inline bool lincomb_shift(uint64_t r[4], const uint64_t a[4], const uint64_t b[4], int64_t f, int64_t g) {
  uint64_t x[5], y[5];
  mul_signed(x, a, f);
  mul_signed(y, b, g);
  u128 c = 0;
  for (int i = 0; i < 5; i++) {
    c += (u128)x[i] + y[i];
    x[i] = (uint64_t)c;
    c >>= 64;
  }
  const bool neg = (x[4] >> 63) != 0;
  if (neg) {
    u128 bb = 1;
    for (int i = 0; i < 5; i++) {
      bb += (uint64_t)~x[i];
      x[i] = (uint64_t)bb;
      bb >>= 64;
    }
  }
  for (int i = 0; i < 4; i++) r[i] = (x[i] >> 31) | (x[i + 1] << 33);
  return neg;
}
// r = (f u + g v) / 2^31 mod m for u, v in [0, m]; signs folded in as f u == |f| (m - u) for f < 0
inline void lincomb_mod(uint64_t r[4], const uint64_t u[4], const uint64_t v[4], int64_t f, int64_t g, const Modulus& M) {
  const uint64_t* src[2] = {u, v};
  const int64_t coef[2] = {f, g};
  uint64_t t[5] = {0, 0, 0, 0, 0};
  for (int k = 0; k < 2; k++) {
    uint64_t w[4];
    const uint64_t mag = coef[k] < 0 ? (uint64_t)0 - (uint64_t)coef[k] : (uint64_t)coef[k];
    if (coef[k] < 0) {  // m - x  (x <= m)
      u128 bw = 0;
      for (int i = 0; i < 4; i++) {
        const u128 d = (u128)M.m[i] - src[k][i] - (uint64_t)bw;
        w[i] = (uint64_t)d;
        bw = (d >> 64) & 1;
      }
    } else {
      memcpy(w, src[k], 32);
    }
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (u128)w[i] * mag + t[i];
      t[i] = (uint64_t)c;
      c >>= 64;
    }
    t[4] += (uint64_t)c;
  }
  // make the low 31 bits vanish with a multiple of m, then divide
  const uint64_t kq = ((t[0] & 0x7fffffffULL) * M.ninv31) & 0x7fffffffULL;
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)M.m[i] * kq + t[i];
    t[i] = (uint64_t)c;
    c >>= 64;
  }
  t[4] += (uint64_t)c;
  uint64_t q[5];
  for (int i = 0; i < 4; i++) q[i] = (t[i] >> 31) | (t[i + 1] << 33);
  q[4] = t[4] >> 31;
  // q < 2m: one conditional subtraction
  uint64_t s[4];
  u128 bw = 0;
  for (int i = 0; i < 4; i++) {
    const u128 d = (u128)q[i] - M.m[i] - (uint64_t)bw;
    s[i] = (uint64_t)d;
    bw = (d >> 64) & 1;
  }
  const bool ge = q[4] != 0 || bw == 0;
  for (int i = 0; i < 4; i++) r[i] = ge ? s[i] : q[i];
}

// out = y^-1 mod m as a canonical residue (0 when y == 0 mod m); y is any integer below 2^256, m an odd prime.
// len(a) + len(b) shrinks by at least 31 bits per round (Pornin, section 3), so 2*256/31 + 1 = 18 rounds bound the
// loop; `false` (never seen: the callers then fall back to the exponentiation) if it has not ended after 40.
inline bool inverse(const uint64_t y[4], const Modulus& M, uint64_t out[4]) {
  uint64_t a[4], b[4], u[4] = {1, 0, 0, 0}, v[4] = {0, 0, 0, 0};
  memcpy(a, y, 32);
  memcpy(b, M.m, 32);
  // invariants: a == u y, b == v y (mod m); a, b >= 0; b odd
  int rounds = 0;
  while (!is_zero4(a)) {
    if (++rounds > 40) return false;
    const int la = bitlen4(a), lb_ = bitlen4(b), n = la > lb_ ? la : lb_;
    uint64_t xa, xb;
    if (n <= 64) {
      xa = a[0];
      xb = b[0];
    } else {
      xa = (bits33_at(a, n - 33) << 31) | (a[0] & 0x7fffffffULL);
      xb = (bits33_at(b, n - 33) << 31) | (b[0] & 0x7fffffffULL);
    }
    uint64_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;  // two's complement in unsigned words
    for (int i = 0; i < 31; i++) {
      const uint64_t odd = (uint64_t)0 - (xa & 1);
      const uint64_t sw = ((uint64_t)0 - (uint64_t)(xa < xb)) & odd;
      uint64_t t = (xa ^ xb) & sw;
      xa ^= t;
      xb ^= t;
      t = (f0 ^ f1) & sw;
      f0 ^= t;
      f1 ^= t;
      t = (g0 ^ g1) & sw;
      g0 ^= t;
      g1 ^= t;
      xa -= xb & odd;
      f0 -= f1 & odd;
      g0 -= g1 & odd;
      xa >>= 1;
      f1 <<= 1;
      g1 <<= 1;
    }
    int64_t F0 = (int64_t)f0, G0 = (int64_t)g0, F1 = (int64_t)f1, G1 = (int64_t)g1;
    uint64_t na[4], nb[4];
    if (lincomb_shift(na, a, b, F0, G0)) {
      F0 = -F0;
      G0 = -G0;
    }
    if (lincomb_shift(nb, a, b, F1, G1)) {
      F1 = -F1;
      G1 = -G1;
    }
    uint64_t nu[4], nv[4];
    lincomb_mod(nu, u, v, F0, G0, M);
    lincomb_mod(nv, u, v, F1, G1, M);
    memcpy(a, na, 32);
    memcpy(b, nb, 32);
    memcpy(u, nu, 32);
    memcpy(v, nv, 32);
  }
  const bool unit = b[0] == 1 && (b[1] | b[2] | b[3]) == 0;
  for (int i = 0; i < 4; i++) out[i] = unit ? v[i] : 0;
  return true;
}

}  // namespace modinv
}  // namespace lb
