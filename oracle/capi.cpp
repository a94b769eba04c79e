// ORACLE — TEST INFRASTRUCTURE ONLY (see field.hpp header).
// C entry points over the oracle so tests/ (ctypes) and bench.py's cpu_baseline /
// --impl reference legs can drive it.  Field elements cross this boundary as 4 x u64
// Montgomery limbs (ark-ff layout); affine points as (x, y) 2 x 4 x u64; extended
// points as (x, y, t, z).
#include <malloc.h>
#include <omp.h>

#include "lasso.hpp"

using namespace oracle;

static inline Fr ldfr(const uint64_t* p) { return Fr::from_raw(p); }
static inline void stfr(uint64_t* p, const Fr& f) { memcpy(p, f.l, 32); }
static inline Fq ldfq(const uint64_t* p) { return Fq::from_raw(p); }
static inline void stfq(uint64_t* p, const Fq& f) { memcpy(p, f.l, 32); }
static inline Affine ldaff(const uint64_t* p) { return Affine{ldfq(p), ldfq(p + 4)}; }
static inline void staff(uint64_t* p, const Affine& a) {
  stfq(p, a.x);
  stfq(p + 4, a.y);
}
static inline Point ldpt(const uint64_t* p) { return Point{ldfq(p), ldfq(p + 4), ldfq(p + 8), ldfq(p + 12)}; }
static inline void stpt(uint64_t* p, const Point& a) {
  stfq(p, a.x);
  stfq(p + 4, a.y);
  stfq(p + 8, a.t);
  stfq(p + 12, a.z);
}
static std::vector<Fr> ldvec(const uint64_t* p, size_t n) {
  std::vector<Fr> v(n);
  for (size_t i = 0; i < n; i++) v[i] = ldfr(p + 4 * i);
  return v;
}

// Some VMs (this build container, possibly the GPU box) back anonymous memory lazily at
// ~20 us per 4 KiB first touch.  Keep freed memory inside the process (no mmap per big
// vector, no trimming) so a warm-up run pays that once and timed runs measure compute.
__attribute__((constructor)) static void orc_malloc_tune() {
  mallopt(M_MMAP_MAX, 0);
  mallopt(M_TRIM_THRESHOLD, -1);
  const char* am = getenv("ORACLE_ARENA_MAX");  // default 1: one shared heap that stays faulted-in
  mallopt(M_ARENA_MAX, am ? atoi(am) : 1);
}

extern "C" {

int orc_num_threads() { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { omp_set_num_threads(n); }

// ---- field (which: 0 = Fr, 1 = Fq) ----
void orc_f_op(int which, int op, const uint64_t* a, const uint64_t* b, uint64_t* out) {
  if (which == 0) {
    Fr x = ldfr(a), y = b ? ldfr(b) : Fr::zero(), r;
    switch (op) {
      case 0: r = x + y; break;
      case 1: r = x - y; break;
      case 2: r = x * y; break;
      case 3: r = x.inverse(); break;
      case 4: r = -x; break;
      default: r = Fr::zero();
    }
    stfr(out, r);
  } else {
    Fq x = ldfq(a), y = b ? ldfq(b) : Fq::zero(), r;
    switch (op) {
      case 0: r = x + y; break;
      case 1: r = x - y; break;
      case 2: r = x * y; break;
      case 3: r = x.inverse(); break;
      case 4: r = -x; break;
      default: r = Fq::zero();
    }
    stfq(out, r);
  }
}
void orc_f_from_u64(int which, uint64_t v, uint64_t* out) {
  if (which == 0) stfr(out, Fr::from_u64(v)); else stfq(out, Fq::from_u64(v));
}
void orc_fr_from_u64_batch(const uint64_t* in, size_t n, uint64_t* out) {
#pragma omp parallel for
  for (size_t i = 0; i < n; i++) stfr(out + 4 * i, Fr::from_u64(in[i]));
}
void orc_f_to_canonical(int which, const uint64_t* a, uint64_t* out) {
  BigInt4 b = which == 0 ? ldfr(a).into_bigint() : ldfq(a).into_bigint();
  memcpy(out, b.l, 32);
}
void orc_f_from_canonical(int which, const uint64_t* a, uint64_t* out) {
  if (which == 0) stfr(out, Fr::from_bigint(a)); else stfq(out, Fq::from_bigint(a));
}
void orc_fr_from_le_bytes_mod_order_64(const uint8_t* in, uint64_t* out) {
  stfr(out, Fr::from_le_bytes_mod_order_64(in));
}

// ---- curve ----
void orc_generator(uint64_t* out_affine) { staff(out_affine, Affine{CC().bx, CC().by}); }
void orc_point_from_affine(const uint64_t* a, uint64_t* out) { stpt(out, Point::from_affine(ldaff(a))); }
void orc_point_add(const uint64_t* a, const uint64_t* b, uint64_t* out) { stpt(out, ldpt(a) + ldpt(b)); }
void orc_point_dbl(const uint64_t* a, uint64_t* out) { stpt(out, ldpt(a).dbl()); }
void orc_point_mul(const uint64_t* a, const uint64_t* s, uint64_t* out) { stpt(out, ldpt(a) * ldfr(s)); }
void orc_point_to_affine(const uint64_t* a, uint64_t* out_affine) { staff(out_affine, ldpt(a).into_affine()); }
void orc_point_compress(const uint64_t* a, uint8_t* out32) { ldpt(a).compress(out32); }
int orc_point_eq(const uint64_t* a, const uint64_t* b) { return ldpt(a) == ldpt(b); }
int orc_decompress(const uint8_t* in32, uint64_t* out_affine) {
  Affine a;
  if (!decompress(in32, a)) return 1;
  staff(out_affine, a);
  return 0;
}
int orc_on_curve(const uint64_t* a) { return on_curve(ldaff(a)); }
// hack: 1 = reference's local msm (small-scalar shortcut), 0 = stock ark-msm behaviour, 2 = naive
int orc_msm(const uint64_t* bases, const uint64_t* scalars, size_t n, int hack, uint64_t* out) {
  std::vector<Affine> B(n);
  std::vector<Fr> S(n);
  for (size_t i = 0; i < n; i++) {
    B[i] = ldaff(bases + 8 * i);
    S[i] = ldfr(scalars + 4 * i);
  }
  Point r;
  if (hack == 2) r = msm_naive(B, S); else msm(B, S, r, hack == 1);
  stpt(out, r);
  return 0;
}
void orc_make_digits(const uint64_t* canonical, size_t w, size_t num_bits, int64_t* out, size_t* count) {
  BigInt4 b;
  memcpy(b.l, canonical, 32);
  auto d = make_digits(b, w, num_bits);
  *count = d.size();
  for (size_t i = 0; i < d.size(); i++) out[i] = d[i];
}
void orc_sample_generators(size_t count, const char* label, uint64_t* out_affine) {
  auto g = sample_generators(count, label);
  for (size_t i = 0; i < count; i++) staff(out_affine + 8 * i, g[i]);
}
// Hyrax row commitments of Z viewed as L_size x R_size (dense_mlpoly.rs:109-128): out = L_size extended points
void orc_commit_rows(const uint64_t* gens_affine /*R_size+1: G.., h*/, const uint64_t* Z, size_t L_size,
                     size_t R_size, uint64_t* out) {
  MultiCommitGens g;
  g.n = R_size;
  for (size_t i = 0; i < R_size; i++) g.G.push_back(Point::from_affine(ldaff(gens_affine + 8 * i)));
  g.h = Point::from_affine(ldaff(gens_affine + 8 * R_size));
  std::vector<Fr> z = ldvec(Z, L_size * R_size);
  Fr zero = Fr::zero();
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t i = 0; i < L_size; i++) stpt(out + 16 * i, batch_commit(&z[R_size * i], R_size, zero, g));
}

// ---- hashes / transcript ----
void orc_keccak_f1600(uint64_t* st) { keccak_f1600(st); }
void orc_shake256(const uint8_t* msg, size_t n, uint8_t* out, size_t outlen) {
  auto o = shake256(std::vector<uint8_t>(msg, msg + n), outlen);
  memcpy(out, o.data(), outlen);
}
void orc_chacha20_words(const uint8_t* seed, uint32_t* out, size_t nwords) {
  ChaCha20Rng rng(seed);
  for (size_t i = 0; i < nwords; i++) out[i] = rng.next_u32();
}
void* orc_transcript_new(const char* label) { return new Transcript(label); }
void orc_transcript_free(void* t) { delete (Transcript*)t; }
void orc_transcript_append_message(void* t, const char* label, const uint8_t* msg, size_t n) {
  ((Transcript*)t)->append_message(label, msg, n);
}
void orc_transcript_append_scalar(void* t, const char* label, const uint64_t* s) {
  ((Transcript*)t)->append_scalar(label, ldfr(s));
}
void orc_transcript_append_point(void* t, const char* label, const uint64_t* p) {
  ((Transcript*)t)->append_point(label, ldpt(p));
}
void orc_transcript_challenge_bytes(void* t, const char* label, uint8_t* out, size_t n) {
  ((Transcript*)t)->challenge_bytes(label, out, n);
}
void orc_transcript_challenge_scalar(void* t, const char* label, uint64_t* out) {
  stfr(out, ((Transcript*)t)->challenge_scalar(label));
}

// ---- polynomials ----
void orc_eq_evals(const uint64_t* r, size_t ell, uint64_t* out) {
  auto ev = EqPolynomial(ldvec(r, ell)).evals();
  memcpy(out, ev.data(), ev.size() * 32);
}
void orc_eq_evaluate(const uint64_t* r, const uint64_t* rx, size_t ell, uint64_t* out) {
  stfr(out, EqPolynomial(ldvec(r, ell)).evaluate(ldvec(rx, ell)));
}
void orc_bind(int top, uint64_t* Z, size_t len, const uint64_t* r) {
  DensePolynomial p(ldvec(Z, len));
  if (top) p.bound_poly_var_top(ldfr(r)); else p.bound_poly_var_bot(ldfr(r));
  memcpy(Z, p.Z.data(), (len / 2) * 32);
}
void orc_evaluate(const uint64_t* Z, size_t len, const uint64_t* r, uint64_t* out) {
  DensePolynomial p(ldvec(Z, len));
  stfr(out, p.evaluate(ldvec(r, p.num_vars)));
}
void orc_bound(const uint64_t* Z, size_t len, const uint64_t* L, uint64_t* out) {
  DensePolynomial p(ldvec(Z, len));
  size_t lv, rv;
  EqPolynomial::compute_factored_lens(p.num_vars, lv, rv);
  auto o = p.bound(ldvec(L, pow2(lv)));
  memcpy(out, o.data(), o.size() * 32);
}
void orc_unipoly_from_evals(const uint64_t* evals, size_t n, uint64_t* coeffs) {
  auto u = UniPoly::from_evals(ldvec(evals, n));
  memcpy(coeffs, u.coeffs.data(), n * 32);
}
void orc_unipoly_evaluate(const uint64_t* coeffs, size_t n, const uint64_t* r, uint64_t* out) {
  UniPoly u{ldvec(coeffs, n)};
  stfr(out, u.evaluate(ldfr(r)));
}
void orc_gaussian_elimination(uint64_t* aug /*n x (n+1)*/, size_t n, uint64_t* out) {
  std::vector<std::vector<Fr>> m(n);
  for (size_t i = 0; i < n; i++) m[i] = ldvec(aug + 4 * i * (n + 1), n + 1);
  auto r = gaussian_elimination(m);
  memcpy(out, r.data(), n * 32);
}

// ---- strategies ----
static Strategy mkS(int kind, size_t C, size_t log_m, size_t log_r) { return Strategy{kind, C, log_m, log_r}; }
size_t orc_num_memories(int kind, size_t C, size_t log_m, size_t log_r) { return mkS(kind, C, log_m, log_r).num_memories(); }
size_t orc_num_subtables(int kind, size_t C, size_t log_m, size_t log_r) { return mkS(kind, C, log_m, log_r).num_subtables(); }
void orc_materialize_subtables(int kind, size_t C, size_t log_m, size_t log_r, uint64_t* out /*nsub x M*/) {
  auto t = mkS(kind, C, log_m, log_r).materialize_subtables();
  size_t M = pow2(log_m);
  for (size_t k = 0; k < t.size(); k++) memcpy(out + 4 * k * M, t[k].data(), M * 32);
}
void orc_evaluate_subtable_mle(int kind, size_t C, size_t log_m, size_t log_r, size_t idx, const uint64_t* point,
                               size_t npoint, uint64_t* out) {
  stfr(out, mkS(kind, C, log_m, log_r).evaluate_subtable_mle(idx, ldvec(point, npoint)));
}
void orc_combine_lookups(int kind, size_t C, size_t log_m, size_t log_r, const uint64_t* vals, uint64_t* out) {
  Strategy S = mkS(kind, C, log_m, log_r);
  auto v = ldvec(vals, S.num_memories());
  stfr(out, S.combine_lookups(v.data()));
}
// gather: E_k[j] = T_{sub(k)}[nz_{dim(k)}[j]]; nz is C x s (u64), out is alpha x s
void orc_lookup_polys(int kind, size_t C, size_t log_m, size_t log_r, const uint64_t* nz, size_t s, uint64_t* out) {
  Strategy S = mkS(kind, C, log_m, log_r);
  std::vector<std::vector<size_t>> idx(C, std::vector<size_t>(s));
  for (size_t i = 0; i < C; i++)
    for (size_t j = 0; j < s; j++) idx[i][j] = nz[i * s + j];
  Subtables st(S, idx, s);
  for (size_t k = 0; k < S.num_memories(); k++) memcpy(out + 4 * k * s, st.lookup_polys[k].Z.data(), s * 32);
}

// one round of the primary sumcheck's evaluation loop (sumcheck.rs:179-237): polys = (alpha+1) x len
void orc_sumcheck_round_arbitrary(int kind, size_t C, size_t log_m, size_t log_r, const uint64_t* polys, size_t len,
                                  uint64_t* evals_out) {
  Strategy S = mkS(kind, C, log_m, log_r);
  size_t alpha = S.num_memories() + 1, deg = S.sumcheck_poly_degree(), half = len / 2;
  std::vector<Fr> ev(deg + 1, Fr::zero()), cur(alpha), nxt(alpha);
  for (size_t i = 0; i < half; i++) {
    for (size_t j = 0; j < alpha; j++) cur[j] = ldfr(polys + 4 * (j * len + i));
    ev[0] += S.combine_lookups_eq(cur.data());
    for (size_t j = 0; j < alpha; j++) cur[j] = ldfr(polys + 4 * (j * len + half + i));
    ev[1] += S.combine_lookups_eq(cur.data());
    for (size_t t = 2; t <= deg; t++) {
      for (size_t j = 0; j < alpha; j++)
        nxt[j] = cur[j] + ldfr(polys + 4 * (j * len + half + i)) - ldfr(polys + 4 * (j * len + i));
      ev[t] += S.combine_lookups_eq(nxt.data());
      cur.swap(nxt);
    }
  }
  memcpy(evals_out, ev.data(), (deg + 1) * 32);
}
// one round of prove_cubic_batched's eval loop (sumcheck.rs:63-89): A, B = ncirc x len; Ceq = len; out = ncirc x 3
void orc_sumcheck_round_cubic(const uint64_t* A, const uint64_t* B, const uint64_t* Ceq, size_t ncirc, size_t len,
                              uint64_t* out) {
  size_t half = len / 2;
  for (size_t k = 0; k < ncirc; k++) {
    Fr p0 = Fr::zero(), p2 = Fr::zero(), p3 = Fr::zero();
    for (size_t i = 0; i < half; i++) {
      Fr a0 = ldfr(A + 4 * (k * len + i)), a1 = ldfr(A + 4 * (k * len + half + i));
      Fr b0 = ldfr(B + 4 * (k * len + i)), b1 = ldfr(B + 4 * (k * len + half + i));
      Fr c0 = ldfr(Ceq + 4 * i), c1 = ldfr(Ceq + 4 * (half + i));
      p0 += a0 * b0 * c0;
      Fr a2 = a1 + a1 - a0, b2 = b1 + b1 - b0, c2 = c1 + c1 - c0;
      p2 += a2 * b2 * c2;
      Fr a3 = a2 + a1 - a0, b3 = b2 + b1 - b0, c3 = c2 + c1 - c0;
      p3 += a3 * b3 * c3;
    }
    stfr(out + 4 * (3 * k), p0);
    stfr(out + 4 * (3 * k + 1), p2);
    stfr(out + 4 * (3 * k + 2), p3);
  }
}

// the reference's sumcheck KAT (sumcheck.rs:459-513): g = product of all polys, scripted challenges.
// polys = npolys x len.  Outputs: round_evals (rounds x (deg+1)), compressed (rounds x deg), final (npolys),
// returns 0 iff the restated verifier returns e == prod(final) and r == scripted.
int orc_sumcheck_product_kat(const uint64_t* polys, size_t npolys, size_t len, const uint64_t* challenges,
                             uint64_t* round_evals, uint64_t* compressed, uint64_t* final_evals,
                             uint64_t* bound_after /*rounds x npolys x (len/2) max; packed per round*/) {
  std::vector<DensePolynomial> P;
  for (size_t j = 0; j < npolys; j++) P.emplace_back(ldvec(polys + 4 * j * len, len));
  size_t rounds = log_2(len), deg = npolys;
  std::vector<Fr> ch = ldvec(challenges, rounds), r, fin;
  Fr claim = Fr::zero();
  for (size_t i = 0; i < len; i++) {
    Fr t = Fr::one();
    for (size_t j = 0; j < npolys; j++) t *= P[j][i];
    claim += t;
  }
  Transcript tp("test_transcript");
  std::vector<std::vector<Fr>> revals;
  auto comb = [&](const Fr* v) {
    Fr t = Fr::one();
    for (size_t j = 0; j < npolys; j++) t *= v[j];
    return t;
  };
  (void)bound_after;
  auto proof = SumcheckInstanceProof::prove_arbitrary(rounds, P, comb, deg, tp, r, fin, &ch, &revals);
  for (size_t k = 0; k < rounds; k++) {
    memcpy(round_evals + 4 * k * (deg + 1), revals[k].data(), (deg + 1) * 32);
    memcpy(compressed + 4 * k * deg, proof.compressed_polys[k].coeffs_except_linear_term.data(), deg * 32);
  }
  memcpy(final_evals, fin.data(), npolys * 32);
  // verifier side with the same scripted challenges: replay decompress/evaluate chain
  Fr e = claim;
  for (size_t k = 0; k < rounds; k++) {
    UniPoly p = UniPoly::decompress(proof.compressed_polys[k], e);
    if (p.degree() != deg) return 2;
    if (p.eval_at_zero() + p.eval_at_one() != e) return 3;
    e = p.evaluate(ch[k]);
  }
  Fr oracle_q = Fr::one();
  for (auto& f : fin) oracle_q *= f;
  return e == oracle_q ? 0 : 1;
}

// grand_product.rs:270-283 fixture: product tree + GP argument prove -> verify
int orc_grand_product_kat(const uint64_t* vals, size_t n, uint64_t* product_out) {
  DensePolynomial p(ldvec(vals, n));
  GrandProductCircuit c(p);
  Fr expected = Fr::one();
  for (size_t i = 0; i < n; i++) expected *= p[i];
  stfr(product_out, c.evaluate());
  if (c.evaluate() != expected) return 1;
  std::vector<GrandProductCircuit*> cs = {&c};
  Transcript tp("test_transcript");
  std::vector<Fr> rand;
  auto proof = BatchedGrandProductArgument::prove(cs, tp, rand);
  Transcript tv("test_transcript");
  std::vector<Fr> claims, rand_v;
  if (!proof.verify({expected}, n, tv, claims, rand_v)) return 2;
  return 0;
}

// dense_mlpoly.rs:586-624 (check_polynomial_commit) and dot_product.rs:350-384 (check_dotproductproof_log) in one:
// commit Z (n = 2^nv elements) with generators sampled from `label`, prove the evaluation at r, verify.
// tamper != 0: the verifier is given eval + 1 and must reject.  eval_out = Z(r).  rc 0 = behaved as expected.
int orc_polyeval_roundtrip(const uint64_t* Z, size_t n, const uint64_t* r, size_t nv, const char* label,
                           const uint64_t* tape_seed, int tamper, uint64_t* eval_out) {
  DensePolynomial poly(ldvec(Z, n));
  if (poly.num_vars != nv) return 10;
  std::vector<Fr> rv = ldvec(r, nv);
  Fr eval = poly.evaluate(rv);
  stfr(eval_out, eval);
  size_t l, rr;
  EqPolynomial::compute_factored_lens(nv, l, rr);
  PolyCommitmentGens gens = PolyCommitmentGens::make(nv, sample_generators(pow2(rr) + 2, label));
  PolyCommitment comm = poly.commit(gens);
  RandomTape tape("proof", ldfr(tape_seed));
  Transcript tp("example");
  PolyEvalProof proof = PolyEvalProof::prove(poly, rv, eval, gens, tp, tape);
  Transcript tv("example");
  Fr claimed = tamper ? eval + Fr::one() : eval;
  bool ok = proof.verify_plain(gens, tv, rv, claimed, comm);
  return (ok == (tamper == 0)) ? 0 : 1;
}

// densified.rs:21-75: indices = n x C (u64, row-major); outputs dim/read (C x s), final (C x m) as integers
void orc_densify(const uint64_t* indices, size_t n, size_t C, size_t log_m, uint64_t* dim, uint64_t* read,
                 uint64_t* fin) {
  size_t s = next_power_of_two(n), m = pow2(log_m);
  for (size_t i = 0; i < C; i++) {
    std::vector<uint64_t> ft(m, 0);
    for (size_t k = 0; k < s; k++) {
      uint64_t addr = k < n ? indices[k * C + i] : 0;
      dim[i * s + k] = addr;
      read[i * s + k] = ft[addr];
      ft[addr]++;
    }
    memcpy(fin + i * m, ft.data(), m * 8);
  }
}
// memory_checking.rs:236-310 fingerprints for one memory; out = init(M) | final(M) | read(s) | write(s)
void orc_gp_fingerprints(const uint64_t* table, size_t M, const uint64_t* dim_usize, const uint64_t* read_ts,
                         const uint64_t* final_ts, size_t s, const uint64_t* gamma, const uint64_t* tau,
                         uint64_t* out) {
  Fr g = ldfr(gamma), t = ldfr(tau), g2 = g.square();
  auto h = [&](const Fr& a, const Fr& v, const Fr& ts) { return ts * g2 + v * g + a - t; };
  for (size_t i = 0; i < M; i++) {
    Fr v = ldfr(table + 4 * i);
    stfr(out + 4 * i, h(Fr::from_u64(i), v, Fr::zero()));
    stfr(out + 4 * (M + i), h(Fr::from_u64(i), v, Fr::from_u64(final_ts[i])));
  }
  for (size_t i = 0; i < s; i++) {
    Fr v = ldfr(table + 4 * dim_usize[i]);
    Fr a = Fr::from_u64(dim_usize[i]), ts = Fr::from_u64(read_ts[i]);
    stfr(out + 4 * (2 * M + i), h(a, v, ts));
    stfr(out + 4 * (2 * M + s + i), h(a, v, ts + Fr::one()));
  }
}

// ---- the whole path: Densify -> commit -> prove (-> verify) ----
// indices: n x C row-major.  gens: affine generator stream of n_gens points (G_0.. ; see surge.rs:32-58).
// flags: bit0 = run verify, bit1 = tamper with the proof before verifying (flip claimed_evaluation),
//        bit2 = tamper a memory-checking element instead, bit3 = skip commit.
// timings_ms: [densify, commit, prove, verify].  Returns 0 ok; 1 verify rejected; <0 error.
int orc_prove(int kind, size_t C, size_t log_m, size_t log_r, const uint64_t* indices, size_t n,
              const uint64_t* r, const uint64_t* gens, size_t n_gens, const uint64_t* tape_seed, int flags,
              uint8_t* proof_out, size_t proof_cap, size_t* proof_len, uint8_t* commit_out, size_t commit_cap,
              size_t* commit_len, uint64_t* challenges_out, size_t challenges_cap, size_t* n_challenges,
              double* timings_ms) {
  try {
    Strategy S = mkS(kind, C, log_m, log_r);
    std::vector<std::vector<size_t>> idx(n, std::vector<size_t>(C));
    for (size_t j = 0; j < n; j++)
      for (size_t i = 0; i < C; i++) idx[j][i] = indices[j * C + i];
    std::vector<Affine> stream(n_gens);
    for (size_t i = 0; i < n_gens; i++) stream[i] = ldaff(gens + 8 * i);
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    spans().ms.clear();
    auto t0 = now();
    DensifiedRepresentation dense = DensifiedRepresentation::from_lookup_indices(idx, C, log_m);
    auto t1 = now();
    if (n_gens < SparsePolyCommitmentGens::needs_points(C, dense.s, S.num_memories(), log_m)) return -2;
    SparsePolyCommitmentGens pg = SparsePolyCommitmentGens::make(C, dense.s, S.num_memories(), log_m, stream);
    auto t1b = now();
    SparsePolynomialCommitment commitment;
    if (!(flags & 8)) commitment = densified_commit(dense, pg);
    auto t2 = now();
    std::vector<Fr> rv = ldvec(r, ark_log2(dense.s));
    RandomTape tape("proof", ldfr(tape_seed));
    Transcript tp("example");
    std::vector<Fr> trace;
    tp.trace = &trace;
    SparsePolynomialEvaluationProof proof = SparsePolynomialEvaluationProof::prove(S, dense, rv, pg, tp, tape);
    auto t3 = now();
    timings_ms[0] = ms(t0, t1);
    timings_ms[1] = ms(t1b, t2);
    timings_ms[2] = ms(t2, t3);
    timings_ms[3] = 0;
    std::vector<uint8_t> pb = serialize_proof(proof);
    *proof_len = pb.size();
    if (proof_out && pb.size() <= proof_cap) memcpy(proof_out, pb.data(), pb.size());
    if (!(flags & 8)) {
      std::vector<uint8_t> cb = serialize_commitment(commitment);
      *commit_len = cb.size();
      if (commit_out && cb.size() <= commit_cap) memcpy(commit_out, cb.data(), cb.size());
    } else {
      *commit_len = 0;
    }
    *n_challenges = trace.size();
    if (challenges_out)
      for (size_t i = 0; i < trace.size() && i < challenges_cap; i++) stfr(challenges_out + 4 * i, trace[i]);
    if ((flags & 1) && !(flags & 8)) {
      if (flags & 2) proof.claimed_evaluation += Fr::one();
      if (flags & 4) proof.memory_check.proof_hash_layer.eval_read[0] += Fr::one();
      Transcript tv("example");
      auto t4 = now();
      bool ok = proof.verify(S, commitment, rv, pg, tv);
      timings_ms[3] = ms(t4, now());
      return ok ? 0 : 1;
    }
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "orc_prove: %s\n", e.what());
    return -1;
  }
}
// span timings of the last orc_prove (the analogue of the reference's tracing log)
size_t orc_spans(char* buf, size_t cap) {
  std::string s;
  for (auto& kv : spans().ms) s += kv.first + "=" + std::to_string(kv.second) + ";";
  if (buf && cap) {
    size_t n = std::min(cap - 1, s.size());
    memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return s.size();
}

}  // extern "C"
