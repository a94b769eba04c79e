// ORACLE — TEST INFRASTRUCTURE ONLY (see field.hpp header).
//
// CPU restatement of (1) the twisted-Edwards group the reference gets from the
// third-party crates ark-ec ^0.4.2 / ark-curve25519 ^0.4.0 (absent from
// /root/reference; "parity unpinned" vs real Rust, pinned here against libsodium
// through pynacl in tests/test_oracle_curve.py), and (2) the reference's in-tree
// Pippenger MSM, /root/reference/src/msm/mod.rs, which IS restated line by line.
#pragma once
#include <algorithm>
#include <cassert>

#include "field.hpp"

namespace oracle {

// -x^2 + y^2 = 1 + d x^2 y^2 over Fq, a = -1, cofactor 8, prime subgroup order l.
struct CurveConsts {
  Fq d, d2, sqrtm1, bx, by;
  CurveConsts() {
    static const uint64_t D[4] = {0x75eb4dca135978a3ULL, 0x00700a4d4141d8abULL,
                                  0x8cc740797779e898ULL, 0x52036cee2b6ffe73ULL};
    static const uint64_t SM1[4] = {0xc4ee1b274a0ea0b0ULL, 0x2f431806ad2fe478ULL,
                                    0x2b4d00993dfbd7a7ULL, 0x2b8324804fc1df0bULL};
    static const uint64_t BX[4] = {0xc9562d608f25d51aULL, 0x692cc7609525a7b2ULL,
                                   0xc0a4e231fdd6dc5cULL, 0x216936d3cd6e53feULL};
    static const uint64_t BY[4] = {0x6666666666666658ULL, 0x6666666666666666ULL,
                                   0x6666666666666666ULL, 0x6666666666666666ULL};
    d = Fq::from_bigint(D);
    d2 = d + d;
    sqrtm1 = Fq::from_bigint(SM1);
    bx = Fq::from_bigint(BX);
    by = Fq::from_bigint(BY);
  }
};
inline const CurveConsts& CC() {
  static CurveConsts c;
  return c;
}

// ark_ec::twisted_edwards::Affine {x, y}: 64 bytes, Montgomery limbs.
struct Affine {
  Fq x, y;
  bool operator==(const Affine& o) const { return x == o.x && y == o.y; }
};

// ark_ec::twisted_edwards::Projective {x, y, t, z} (extended coordinates), 128 bytes.
struct Point {
  Fq x, y, t, z;

  static Point zero() { return Point{Fq::zero(), Fq::one(), Fq::zero(), Fq::one()}; }
  static Point from_affine(const Affine& a) { return Point{a.x, a.y, a.x * a.y, Fq::one()}; }
  static Point generator() { return from_affine(Affine{CC().bx, CC().by}); }
  bool is_zero() const { return x.is_zero() && y == z; }

  // unified addition add-2008-hwcd-3 (a = -1); any complete formula gives the same group element
  Point operator+(const Point& o) const {
    Fq A = (y - x) * (o.y - o.x);
    Fq B = (y + x) * (o.y + o.x);
    Fq C = t * CC().d2 * o.t;
    Fq D = (z * o.z).dbl();
    Fq E = B - A, F = D - C, G = D + C, H = B + A;
    return Point{E * F, G * H, E * H, F * G};
  }
  Point add_affine(const Affine& o) const {
    Fq ot = o.x * o.y;
    Fq A = (y - x) * (o.y - o.x);
    Fq B = (y + x) * (o.y + o.x);
    Fq C = t * CC().d2 * ot;
    Fq D = z.dbl();
    Fq E = B - A, F = D - C, G = D + C, H = B + A;
    return Point{E * F, G * H, E * H, F * G};
  }
  Point sub_affine(const Affine& o) const { return add_affine(Affine{-o.x, o.y}); }
  Point dbl() const {
    Fq A = x.square(), B = y.square(), C = z.square().dbl();
    Fq D = -A;
    Fq E = (x + y).square() - A - B;
    Fq G = D + B, F = G - C, H = D - B;
    return Point{E * F, G * H, E * H, F * G};
  }
  Point operator-() const { return Point{-x, y, -t, z}; }
  Point& operator+=(const Point& o) { return *this = *this + o; }

  // scalar multiplication by a canonical integer (double-and-add, MSB first)
  Point mul_bigint(const BigInt4& k) const {
    Point acc = zero();
    bool started = false;
    for (int i = 255; i >= 0; i--) {
      if (started) acc = acc.dbl();
      if ((k.l[i / 64] >> (i % 64)) & 1) {
        acc = acc + *this;
        started = true;
      }
    }
    return acc;
  }
  Point operator*(const Fr& s) const { return mul_bigint(s.into_bigint()); }

  Affine into_affine() const {
    Fq zi = z.inverse();
    return Affine{x * zi, y * zi};
  }
  // projective equality
  bool operator==(const Point& o) const { return x * o.z == o.x * z && y * o.z == o.y * z; }
  bool operator!=(const Point& o) const { return !(*this == o); }

  // ark-serialize compressed TE point: 32-byte LE y, top bit of last byte set iff
  // x is "negative" in arkworks' sense (x > -x as canonical integers).  [memory; SURVEY App. C]
  void compress(uint8_t out[32]) const {
    Affine a = into_affine();
    a.y.to_bytes(out);
    Fq nx = -a.x;
    if (a.x.canonical_gt(nx)) out[31] |= 0x80;
  }
};

inline void compress_affine(const Affine& a, uint8_t out[32]) {
  a.y.to_bytes(out);
  Fq nx = -a.x;
  if (a.x.canonical_gt(nx)) out[31] |= 0x80;
}

// CurveGroup::normalize_batch — Montgomery batch inversion
inline std::vector<Affine> normalize_batch(const std::vector<Point>& v) {
  size_t n = v.size();
  std::vector<Fq> pref(n);
  Fq acc = Fq::one();
  for (size_t i = 0; i < n; i++) {
    pref[i] = acc;
    acc = acc * v[i].z;
  }
  Fq inv = acc.inverse();
  std::vector<Affine> out(n);
  for (size_t i = n; i-- > 0;) {
    Fq zi = inv * pref[i];
    inv = inv * v[i].z;
    out[i] = Affine{v[i].x * zi, v[i].y * zi};
  }
  return out;
}

// sqrt in Fq (q = 5 mod 8); returns false if no root
inline bool fq_sqrt(const Fq& a, Fq& out) {
  // candidate = a^((q+3)/8)
  static const uint64_t E[4] = {0xfffffffffffffffeULL, 0xffffffffffffffffULL, 0xffffffffffffffffULL,
                                0x0fffffffffffffffULL};  // (q+3)/8 = 2^252 - 2
  Fq c = a.pow(E);
  if (c.square() == a) {
    out = c;
    return true;
  }
  c = c * CC().sqrtm1;
  if (c.square() == a) {
    out = c;
    return true;
  }
  return false;
}

// Affine::get_point_from_y_unchecked(y, greatest)  [ark-ec, memory]: x^2 = (y^2-1)/(d y^2 + 1)
inline bool point_from_y(const Fq& y, bool greatest, Affine& out) {
  Fq y2 = y.square();
  Fq num = y2 - Fq::one();
  Fq den = CC().d * y2 + Fq::one();
  if (den.is_zero()) return false;
  Fq x2 = num * den.inverse();
  Fq x;
  if (!fq_sqrt(x2, x)) return false;
  Fq nx = -x;
  bool x_is_greater = x.canonical_gt(nx);
  out = Affine{(x_is_greater == greatest) ? x : nx, y};
  return true;
}
inline bool decompress(const uint8_t in[32], Affine& out) {
  uint8_t b[32];
  memcpy(b, in, 32);
  bool neg = (b[31] & 0x80) != 0;
  b[31] &= 0x7f;
  uint64_t raw[4];
  memcpy(raw, b, 32);
  if (Fq::geq_mod(raw)) return false;
  return point_from_y(Fq::from_bigint(raw), neg, out);
}
inline bool on_curve(const Affine& a) {
  Fq x2 = a.x.square(), y2 = a.y.square();
  return y2 - x2 == Fq::one() + CC().d * x2 * y2;
}

// ---------------------------------------------------------------------------------------
// /root/reference/src/msm/mod.rs restated.
// ---------------------------------------------------------------------------------------

// ark_std::log2 — ceil(log2 x), 0 for x <= 1
inline uint32_t ark_log2(size_t x) {
  if (x <= 1) return 0;
  return 64 - __builtin_clzll((unsigned long long)(x - 1));
}
// msm/mod.rs:322-325
inline size_t ln_without_floats(size_t a) { return (size_t)(ark_log2(a) * 69 / 100); }

// msm/mod.rs:277-316 make_digits
inline std::vector<int64_t> make_digits(const BigInt4& a, size_t w, size_t num_bits) {
  const uint64_t* scalar = a.l;
  uint64_t radix = 1ULL << w;
  uint64_t window_mask = radix - 1;
  uint64_t carry = 0;
  if (num_bits == 0) num_bits = a.num_bits();
  size_t digits_count = (num_bits + w - 1) / w;
  std::vector<int64_t> digits(digits_count, 0);
  for (size_t i = 0; i < digits_count; i++) {
    size_t bit_offset = i * w;
    size_t u64_idx = bit_offset / 64;
    size_t bit_idx = bit_offset % 64;
    uint64_t bit_buf;
    if (bit_idx < 64 - w || u64_idx == 3) {
      bit_buf = scalar[u64_idx] >> bit_idx;
    } else {
      bit_buf = (scalar[u64_idx] >> bit_idx) | (scalar[1 + u64_idx] << (64 - bit_idx));
    }
    uint64_t coef = carry + (bit_buf & window_mask);
    carry = (coef + radix / 2) >> w;
    digits[i] = (int64_t)coef - (int64_t)(carry << w);
  }
  digits[digits_count - 1] += (int64_t)(carry << w);
  return digits;
}

// msm/mod.rs:91-164 msm_bigint_wnaf.  `small_scalar_hack` = false reproduces what
// `--features ark-msm` (stock ark_ec::VariableBaseMSM) computes: num_bits = 253 always.
inline Point msm_bigint_wnaf(const Affine* bases, const BigInt4* bigints, size_t n,
                             bool small_scalar_hack = true) {
  size_t max_num_bits = 1;
  if (small_scalar_hack) {
    for (size_t i = 0; i < n; i++) {
      if (bigints[i].num_bits() > max_num_bits) max_num_bits = bigints[i].num_bits();
      if (max_num_bits > 60) {
        max_num_bits = FrParams::MODULUS_BIT_SIZE;
        break;
      }
    }
  } else {
    max_num_bits = FrParams::MODULUS_BIT_SIZE;
  }
  size_t size = n;
  size_t c = size < 32 ? 3 : ln_without_floats(size) + 2;
  size_t num_bits = max_num_bits;
  size_t digits_count = (num_bits + c - 1) / c;
  std::vector<int64_t> scalar_digits;
  scalar_digits.reserve(size * digits_count);
  for (size_t i = 0; i < size; i++) {
    std::vector<int64_t> d = make_digits(bigints[i], c, num_bits);
    scalar_digits.insert(scalar_digits.end(), d.begin(), d.end());
  }
  Point zero = Point::zero();
  std::vector<Point> window_sums(digits_count);
  // cfg_into_iter!: sequential here — the reference crate defines no "parallel" feature
  // for this module, so windows run one after another.
  for (size_t i = 0; i < digits_count; i++) {
    std::vector<Point> buckets((size_t)1 << c, zero);
    for (size_t k = 0; k < size; k++) {
      int64_t scalar = scalar_digits[k * digits_count + i];
      if (scalar > 0)
        buckets[(size_t)(scalar - 1)] = buckets[(size_t)(scalar - 1)].add_affine(bases[k]);
      else if (scalar < 0)
        buckets[(size_t)(-scalar - 1)] = buckets[(size_t)(-scalar - 1)].sub_affine(bases[k]);
    }
    Point running_sum = zero, res = zero;
    for (size_t b = buckets.size(); b-- > 0;) {
      running_sum += buckets[b];
      res += running_sum;
    }
    window_sums[i] = res;
  }
  Point lowest = window_sums[0];
  Point total = zero;
  for (size_t i = digits_count; i-- > 1;) {
    total += window_sums[i];
    for (size_t k = 0; k < c; k++) total = total.dbl();
  }
  return lowest + total;
}

// VariableBaseMSM::msm (msm/mod.rs:36-40) + msm_unchecked (22-27).  Returns false on
// length mismatch (the reference's Err(min_len)); every caller unwraps.
inline bool msm(const std::vector<Affine>& bases, const std::vector<Fr>& scalars, Point& out,
                bool small_scalar_hack = true) {
  if (bases.size() != scalars.size()) return false;
  std::vector<BigInt4> bigints(scalars.size());
  for (size_t i = 0; i < scalars.size(); i++) bigints[i] = scalars[i].into_bigint();
  out = msm_bigint_wnaf(bases.data(), bigints.data(), bases.size(), small_scalar_hack);
  return true;
}

// naive sum_i s_i * B_i, used only to cross-check the Pippenger restatement
inline Point msm_naive(const std::vector<Affine>& bases, const std::vector<Fr>& scalars) {
  Point acc = Point::zero();
  for (size_t i = 0; i < bases.size(); i++) acc += Point::from_affine(bases[i]) * scalars[i];
  return acc;
}

}  // namespace oracle
