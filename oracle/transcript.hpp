// ORACLE — TEST INFRASTRUCTURE ONLY (see field.hpp header).
//
// Restatement of the third-party transcript stack the reference uses through
// /root/reference/src/utils/transcript.rs and utils/random.rs:
//   merlin ^3.0.0 (STROBE-128 over Keccak-f[1600]) — published algorithm, pinned by the
//   Merlin test vector in tests/test_oracle_transcript.py;
//   sha3 ^0.8.2 Shake256 — pinned against hashlib.shake_256;
//   rand_chacha ^0.3.0 ChaCha20Rng — pinned against RFC 8439 block-function vector.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "curve.hpp"

namespace oracle {

inline uint64_t rotl64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

inline void keccak_f1600(uint64_t st[25]) {
  static const uint64_t RC[24] = {
      0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
      0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
      0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
      0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
      0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
      0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  static const int ROT[25] = {0,  1,  62, 28, 27, 36, 44, 6,  55, 20, 3,  10, 43,
                              25, 39, 41, 45, 15, 21, 8,  18, 2,  61, 56, 14};
  for (int round = 0; round < 24; round++) {
    uint64_t C[5], D[5], B[25];
    for (int x = 0; x < 5; x++) C[x] = st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20];
    for (int x = 0; x < 5; x++) D[x] = C[(x + 4) % 5] ^ rotl64(C[(x + 1) % 5], 1);
    for (int i = 0; i < 25; i++) st[i] ^= D[i % 5];
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) B[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(st[x + 5 * y], ROT[x + 5 * y]);
    for (int y = 0; y < 5; y++)
      for (int x = 0; x < 5; x++)
        st[x + 5 * y] = B[x + 5 * y] ^ ((~B[(x + 1) % 5 + 5 * y]) & B[(x + 2) % 5 + 5 * y]);
    st[0] ^= RC[round];
  }
}

// Shake256 XOF (sponge rate 136, domain 0x1f); commitments.rs:23-31
inline std::vector<uint8_t> shake256(const std::vector<uint8_t>& msg, size_t outlen) {
  const size_t rate = 136;
  uint8_t st[200];
  memset(st, 0, 200);
  size_t pos = 0;
  auto permute = [&]() {
    uint64_t w[25];
    memcpy(w, st, 200);
    keccak_f1600(w);
    memcpy(st, w, 200);
  };
  for (uint8_t b : msg) {
    st[pos++] ^= b;
    if (pos == rate) {
      permute();
      pos = 0;
    }
  }
  st[pos] ^= 0x1f;
  st[rate - 1] ^= 0x80;
  permute();
  std::vector<uint8_t> out;
  pos = 0;
  while (out.size() < outlen) {
    if (pos == rate) {
      permute();
      pos = 0;
    }
    out.push_back(st[pos++]);
  }
  return out;
}

// merlin::strobe::Strobe128
struct Strobe128 {
  static constexpr int R = 166;
  static constexpr uint8_t FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32;
  uint8_t st[200];
  uint8_t pos, pos_begin, cur_flags;

  explicit Strobe128(const char* protocol_label) {
    memset(st, 0, 200);
    const uint8_t init[6] = {1, R + 2, 1, 0, 1, 96};
    memcpy(st, init, 6);
    memcpy(st + 6, "STROBEv1.0.2", 12);
    permute();
    pos = 0;
    pos_begin = 0;
    cur_flags = 0;
    meta_ad((const uint8_t*)protocol_label, strlen(protocol_label), false);
  }
  void permute() {
    uint64_t w[25];
    memcpy(w, st, 200);
    keccak_f1600(w);
    memcpy(st, w, 200);
  }
  void run_f() {
    st[pos] ^= pos_begin;
    st[pos + 1] ^= 0x04;
    st[R + 1] ^= 0x80;
    permute();
    pos = 0;
    pos_begin = 0;
  }
  void absorb(const uint8_t* d, size_t n) {
    for (size_t i = 0; i < n; i++) {
      st[pos] ^= d[i];
      pos++;
      if (pos == R) run_f();
    }
  }
  void squeeze(uint8_t* d, size_t n) {
    for (size_t i = 0; i < n; i++) {
      d[i] = st[pos];
      st[pos] = 0;
      pos++;
      if (pos == R) run_f();
    }
  }
  void begin_op(uint8_t flags, bool more) {
    if (more) {
      assert(cur_flags == flags);
      return;
    }
    assert((flags & FLAG_T) == 0);
    uint8_t old_begin = pos_begin;
    pos_begin = pos + 1;
    cur_flags = flags;
    uint8_t hdr[2] = {old_begin, flags};
    absorb(hdr, 2);
    bool force_f = (flags & (FLAG_C | FLAG_K)) != 0;
    if (force_f && pos != 0) run_f();
  }
  void meta_ad(const uint8_t* d, size_t n, bool more) {
    begin_op(FLAG_M | FLAG_A, more);
    absorb(d, n);
  }
  void ad(const uint8_t* d, size_t n, bool more) {
    begin_op(FLAG_A, more);
    absorb(d, n);
  }
  void prf(uint8_t* d, size_t n, bool more) {
    begin_op(FLAG_I | FLAG_A | FLAG_C, more);
    squeeze(d, n);
  }
};

// merlin::Transcript + the reference's ProofTranscript impl (utils/transcript.rs:20-72)
struct Transcript {
  Strobe128 strobe;
  std::vector<Fr>* trace = nullptr;  // test hook: every challenge scalar, in order
  explicit Transcript(const char* label) : strobe("Merlin v1.0") {
    append_message("dom-sep", (const uint8_t*)label, strlen(label));
  }
  void append_message(const char* label, const uint8_t* msg, size_t n) {
    uint32_t len = (uint32_t)n;
    uint8_t le[4] = {(uint8_t)len, (uint8_t)(len >> 8), (uint8_t)(len >> 16), (uint8_t)(len >> 24)};
    strobe.meta_ad((const uint8_t*)label, strlen(label), false);
    strobe.meta_ad(le, 4, true);
    strobe.ad(msg, n, false);
  }
  void append_message(const char* label, const char* msg) {
    append_message(label, (const uint8_t*)msg, strlen(msg));
  }
  void append_u64(const char* label, uint64_t x) {
    uint8_t b[8];
    memcpy(b, &x, 8);
    append_message(label, b, 8);
  }
  void challenge_bytes(const char* label, uint8_t* dest, size_t n) {
    uint32_t len = (uint32_t)n;
    uint8_t le[4] = {(uint8_t)len, (uint8_t)(len >> 8), (uint8_t)(len >> 16), (uint8_t)(len >> 24)};
    strobe.meta_ad((const uint8_t*)label, strlen(label), false);
    strobe.meta_ad(le, 4, true);
    strobe.prf(dest, n, false);
  }
  // ---- ProofTranscript (utils/transcript.rs) ----
  void append_protocol_name(const char* name) { append_message("protocol-name", name); }
  void append_scalar(const char* label, const Fr& s) {
    uint8_t b[32];
    s.to_bytes(b);
    append_message(label, b, 32);
  }
  void append_scalars(const char* label, const std::vector<Fr>& v) {
    append_message(label, "begin_append_vector");
    for (const Fr& s : v) append_scalar(label, s);
    append_message(label, "end_append_vector");
  }
  void append_point(const char* label, const Point& p) {
    uint8_t b[32];
    p.compress(b);
    append_message(label, b, 32);
  }
  Fr challenge_scalar(const char* label) {
    uint8_t buf[64];
    challenge_bytes(label, buf, 64);
    Fr c = Fr::from_le_bytes_mod_order_64(buf);
    if (trace) trace->push_back(c);
    return c;
  }
  std::vector<Fr> challenge_vector(const char* label, size_t len) {
    std::vector<Fr> v;
    for (size_t i = 0; i < len; i++) v.push_back(challenge_scalar(label));
    return v;
  }
};

// utils/random.rs:9-39.  The reference seeds the tape with F::rand(test_rng()); that RNG
// lives in ark-std (absent), so the seed scalar is an explicit input of the parity
// contract.
struct RandomTape {
  Transcript tape;
  RandomTape(const char* name, const Fr& init_randomness) : tape(name) {
    tape.append_scalar("init_randomness", init_randomness);
  }
  Fr random_scalar(const char* label) { return tape.challenge_scalar(label); }
  std::vector<Fr> random_vector(const char* label, size_t len) {
    return tape.challenge_vector(label, len);
  }
};

// rand_chacha::ChaCha20Rng::from_seed — RFC 8439 block function with a 64-bit block
// counter and zero stream id, words consumed in order.
struct ChaCha20Rng {
  uint32_t key[8];
  uint64_t counter = 0;
  uint32_t buf[16];
  int idx = 16;
  explicit ChaCha20Rng(const uint8_t seed[32]) { memcpy(key, seed, 32); }
  static uint32_t rotl(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
  static void qr(uint32_t* s, int a, int b, int c, int d) {
    s[a] += s[b]; s[d] ^= s[a]; s[d] = rotl(s[d], 16);
    s[c] += s[d]; s[b] ^= s[c]; s[b] = rotl(s[b], 12);
    s[a] += s[b]; s[d] ^= s[a]; s[d] = rotl(s[d], 8);
    s[c] += s[d]; s[b] ^= s[c]; s[b] = rotl(s[b], 7);
  }
  void refill() {
    uint32_t s[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574};
    for (int i = 0; i < 8; i++) s[4 + i] = key[i];
    s[12] = (uint32_t)counter;
    s[13] = (uint32_t)(counter >> 32);
    s[14] = 0;
    s[15] = 0;
    uint32_t w[16];
    memcpy(w, s, 64);
    for (int r = 0; r < 10; r++) {
      qr(w, 0, 4, 8, 12); qr(w, 1, 5, 9, 13); qr(w, 2, 6, 10, 14); qr(w, 3, 7, 11, 15);
      qr(w, 0, 5, 10, 15); qr(w, 1, 6, 11, 12); qr(w, 2, 7, 8, 13); qr(w, 3, 4, 9, 14);
    }
    for (int i = 0; i < 16; i++) buf[i] = w[i] + s[i];
    counter++;
    idx = 0;
  }
  uint32_t next_u32() {
    if (idx == 16) refill();
    return buf[idx++];
  }
  uint64_t next_u64() {
    uint64_t lo = next_u32();
    uint64_t hi = next_u32();
    return lo | (hi << 32);
  }
};

// Fp::rand [ark-ff, memory]: 4 x next_u64, mask to the modulus bit-length, reject if >= p,
// the accepted bits ARE the Montgomery representation.
template <class P>
inline Fp<P> fp_rand(ChaCha20Rng& rng) {
  for (;;) {
    uint64_t raw[4];
    for (int i = 0; i < 4; i++) raw[i] = rng.next_u64();
    raw[3] &= (~0ULL) >> (256 - P::MODULUS_BIT_SIZE);
    if (!Fp<P>::geq_mod(raw)) return Fp<P>::from_raw(raw);
  }
}

// TE Projective::rand [ark-ec, memory]: y <- Fq::rand; greatest <- bool; point from y; clear cofactor.
inline Point point_rand(ChaCha20Rng& rng) {
  for (;;) {
    Fq y = fp_rand<FqParams>(rng);
    bool greatest = (int32_t)rng.next_u32() < 0;
    Affine a;
    if (point_from_y(y, greatest, a)) {
      Point p = Point::from_affine(a);
      return p.dbl().dbl().dbl();
    }
  }
}

}  // namespace oracle
