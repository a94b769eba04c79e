// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// build, link or execute anything under oracle/.
//
// CPU restatement of the prime-field arithmetic the reference gets from the
// third-party crate ark-ff ^0.4.2 (`Fp<MontBackend<_,4>,4>`; NOT under
// /root/reference — see SURVEY.md §8c).  "parity unpinned" at the byte level
// against real Rust; pinned here against Python big-ints (tests/test_oracle_field.py)
// and the small-integer KATs of the reference's own tests.
//
// Layout: 4 x u64 little-endian limbs holding a*R mod p, R = 2^256 (ark-ff
// Montgomery form) -> a `&[Fr]` in Rust is bit-identical to an array of these.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace oracle {

typedef unsigned __int128 u128;

struct FrParams {
  // l = 2^252 + 27742317777372353535851937790883648493 (curve25519 scalar field)
  static constexpr uint64_t MOD[4] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0x0ULL,
                                      0x1000000000000000ULL};
  static constexpr uint64_t R[4] = {0xd6ec31748d98951dULL, 0xc6ef5bf4737dcf70ULL,
                                    0xfffffffffffffffeULL, 0x0fffffffffffffffULL};
  static constexpr uint64_t R2[4] = {0xa40611e3449c0f01ULL, 0xd00e1ba768859347ULL,
                                     0xceec73d217f5be65ULL, 0x0399411b7c309a3dULL};
  static constexpr uint64_t INV = 0xd2b51da312547e1bULL;  // -l^-1 mod 2^64
  static constexpr int MODULUS_BIT_SIZE = 253;
};

struct FqParams {
  // q = 2^255 - 19 (curve25519 base field)
  static constexpr uint64_t MOD[4] = {0xffffffffffffffedULL, 0xffffffffffffffffULL,
                                      0xffffffffffffffffULL, 0x7fffffffffffffffULL};
  static constexpr uint64_t R[4] = {38, 0, 0, 0};
  static constexpr uint64_t R2[4] = {1444, 0, 0, 0};
  static constexpr uint64_t INV = 0x86bca1af286bca1bULL;
  static constexpr int MODULUS_BIT_SIZE = 255;
};

// 256-bit little-endian integer, the analogue of ark_ff::BigInt<4>.
struct BigInt4 {
  uint64_t l[4];
  bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
  // ark_ff::BigInteger::num_bits
  uint32_t num_bits() const {
    for (int i = 3; i >= 0; i--)
      if (l[i]) return 64 * i + (64 - __builtin_clzll(l[i]));
    return 0;
  }
};

template <class P>
struct Fp {
  uint64_t l[4];

  static Fp zero() { return Fp{{0, 0, 0, 0}}; }
  static Fp one() { return Fp{{P::R[0], P::R[1], P::R[2], P::R[3]}}; }
  static Fp from_raw(const uint64_t r[4]) { return Fp{{r[0], r[1], r[2], r[3]}}; }

  bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
  bool operator==(const Fp& o) const {
    return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3];
  }
  bool operator!=(const Fp& o) const { return !(*this == o); }

  static bool geq_mod(const uint64_t a[4]) {
    for (int i = 3; i >= 0; i--) {
      if (a[i] > P::MOD[i]) return true;
      if (a[i] < P::MOD[i]) return false;
    }
    return true;
  }
  static void sub_mod(uint64_t a[4]) {
    u128 b = 0;
    for (int i = 0; i < 4; i++) {
      u128 d = (u128)a[i] - P::MOD[i] - (uint64_t)b;
      a[i] = (uint64_t)d;
      b = (d >> 64) & 1;
    }
  }

  Fp operator+(const Fp& o) const {
    Fp r;
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (u128)l[i] + o.l[i];
      r.l[i] = (uint64_t)c;
      c >>= 64;
    }
    // both moduli are < 2^255 so no carry out of 256 bits
    if (geq_mod(r.l)) sub_mod(r.l);
    return r;
  }
  Fp operator-(const Fp& o) const {
    Fp r;
    u128 b = 0;
    for (int i = 0; i < 4; i++) {
      u128 d = (u128)l[i] - o.l[i] - (uint64_t)b;
      r.l[i] = (uint64_t)d;
      b = (d >> 64) & 1;
    }
    if (b) {
      u128 c = 0;
      for (int i = 0; i < 4; i++) {
        c += (u128)r.l[i] + P::MOD[i];
        r.l[i] = (uint64_t)c;
        c >>= 64;
      }
    }
    return r;
  }
  Fp operator-() const { return is_zero() ? *this : zero() - *this; }
  Fp& operator+=(const Fp& o) { return *this = *this + o; }
  Fp& operator-=(const Fp& o) { return *this = *this - o; }
  Fp& operator*=(const Fp& o) { return *this = *this * o; }

  // CIOS Montgomery multiplication: a*b*R^-1 mod p
  static void mont_mul(uint64_t out[4], const uint64_t a[4], const uint64_t b[4]) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
      u128 c = 0;
      for (int j = 0; j < 4; j++) {
        c += (u128)t[j] + (u128)a[j] * b[i];
        t[j] = (uint64_t)c;
        c >>= 64;
      }
      c += t[4];
      t[4] = (uint64_t)c;
      t[5] = (uint64_t)(c >> 64);
      uint64_t m = t[0] * P::INV;
      c = (u128)t[0] + (u128)m * P::MOD[0];
      c >>= 64;
      for (int j = 1; j < 4; j++) {
        c += (u128)t[j] + (u128)m * P::MOD[j];
        t[j - 1] = (uint64_t)c;
        c >>= 64;
      }
      c += t[4];
      t[3] = (uint64_t)c;
      t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || geq_mod(t)) sub_mod(t);
    out[0] = t[0];
    out[1] = t[1];
    out[2] = t[2];
    out[3] = t[3];
  }
  Fp operator*(const Fp& o) const {
    Fp r;
    mont_mul(r.l, l, o.l);
    return r;
  }
  Fp square() const { return *this * *this; }
  Fp dbl() const { return *this + *this; }

  // F::from(u64)
  static Fp from_u64(uint64_t v) {
    uint64_t raw[4] = {v, 0, 0, 0};
    Fp r;
    mont_mul(r.l, raw, P::R2);
    return r;
  }
  // PrimeField::into_bigint — canonical integer
  BigInt4 into_bigint() const {
    static const uint64_t ONE[4] = {1, 0, 0, 0};
    BigInt4 r;
    mont_mul(r.l, l, ONE);
    return r;
  }
  // from a canonical (or any < 2^256) integer
  static Fp from_bigint(const uint64_t raw[4]) {
    Fp r;
    mont_mul(r.l, raw, P::R2);
    return r;
  }
  // ark_ff serialize_compressed for a field element: 32 bytes LE canonical
  void to_bytes(uint8_t out[32]) const {
    BigInt4 b = into_bigint();
    memcpy(out, b.l, 32);
  }
  static Fp from_bytes32_mod_order(const uint8_t in[32]) {
    uint64_t raw[4];
    memcpy(raw, in, 32);
    return from_bigint(raw);
  }
  // PrimeField::from_le_bytes_mod_order on 64 bytes (utils/transcript.rs:61-65)
  static Fp from_le_bytes_mod_order_64(const uint8_t in[64]) {
    uint64_t lo[4], hi[4];
    memcpy(lo, in, 32);
    memcpy(hi, in + 32, 32);
    Fp a = from_bigint(lo);           // lo mod p        (Montgomery form)
    Fp b = from_bigint(hi);           // hi mod p
    Fp r2 = from_raw(P::R2);          // Montgomery form of R = 2^256
    return a + b * r2;                // lo + hi * 2^256
  }

  Fp pow(const uint64_t e[4]) const {
    Fp acc = one();
    for (int i = 255; i >= 0; i--) {
      acc = acc.square();
      if ((e[i / 64] >> (i % 64)) & 1) acc = acc * *this;
    }
    return acc;
  }
  // Field::inverse via Fermat (value is unique, algorithm is irrelevant for parity)
  Fp inverse() const {
    uint64_t e[4] = {P::MOD[0] - 2, P::MOD[1], P::MOD[2], P::MOD[3]};
    return pow(e);
  }
  // canonical-integer comparison, used for the TE sign flag
  bool canonical_gt(const Fp& o) const {
    BigInt4 a = into_bigint(), b = o.into_bigint();
    for (int i = 3; i >= 0; i--) {
      if (a.l[i] > b.l[i]) return true;
      if (a.l[i] < b.l[i]) return false;
    }
    return false;
  }
  std::string hex() const {
    uint8_t b[32];
    to_bytes(b);
    static const char* H = "0123456789abcdef";
    std::string s;
    for (int i = 31; i >= 0; i--) {
      s.push_back(H[b[i] >> 4]);
      s.push_back(H[b[i] & 15]);
    }
    return s;
  }
};

typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

}  // namespace oracle
