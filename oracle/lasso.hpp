// ORACLE — TEST INFRASTRUCTURE ONLY (see field.hpp header).
//
// CPU restatement of the reference's Lasso prover + verifier, following (file:line
// relative to /root/reference/src):
//   poly/{dense_mlpoly,eq_poly,unipoly,commitments,identity_poly}.rs
//   subprotocols/{sumcheck,grand_product,dot_product,bullet}.rs
//   subtables/{mod,and,or,xor,lt,range_check}.rs
//   lasso/{densified,surge,memory_checking}.rs, utils/{math,mod,gaussian_elimination}.rs
// The reference is Rust and cannot be built here (no cargo/rustc, crates not vendored),
// so this is a "port"; it is threaded (OpenMP) on the same axes the reference uses rayon
// and serial where the reference is serial, because it is also the timed CPU baseline.
// PARITY UNPINNED against the real Rust binary at the byte level; pinned against every
// known-answer test the reference holds for this path (tests/test_oracle_*.py).
#pragma once
#include <array>
#include <chrono>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>

#include "transcript.hpp"

namespace oracle {

// ---------------------------------------------------------------- utils/math.rs
inline size_t pow2(size_t n) { return (size_t)1 << n; }
// Math::log_2 (utils/math.rs:27-35): exact for powers of two, ceil otherwise
inline size_t log_2(size_t x) {
  assert(x != 0);
  if ((x & (x - 1)) == 0) return (size_t)__builtin_ctzll((unsigned long long)x);
  return 64 - (size_t)__builtin_clzll((unsigned long long)x);
}
inline size_t next_power_of_two(size_t x) {
  size_t p = 1;
  while (p < x) p <<= 1;
  return p;
}
// utils/mod.rs:82-89
inline void split_bits(size_t item, size_t num_bits, size_t& high, size_t& low) {
  size_t max_value = ((size_t)1 << num_bits) - 1;
  low = item & max_value;
  high = (item >> num_bits) & max_value;
}

struct Spans {  // the analogue of the reference's tracing spans
  std::map<std::string, double> ms;
  void add(const std::string& k, double v) { ms[k] += v; }
};
inline Spans& spans() {
  static Spans s;
  return s;
}
struct SpanTimer {
  std::string name;
  std::chrono::steady_clock::time_point t0;
  explicit SpanTimer(const char* n) : name(n), t0(std::chrono::steady_clock::now()) {}
  ~SpanTimer() {
    spans().add(name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  }
};

// ---------------------------------------------------------------- poly/eq_poly.rs
struct EqPolynomial {
  std::vector<Fr> r;
  explicit EqPolynomial(std::vector<Fr> r_) : r(std::move(r_)) {}
  // eq_poly.rs:14-19
  Fr evaluate(const std::vector<Fr>& rx) const {
    assert(r.size() == rx.size());
    Fr acc = Fr::one();
    for (size_t i = 0; i < rx.size(); i++)
      acc = acc * (r[i] * rx[i] + (Fr::one() - r[i]) * (Fr::one() - rx[i]));
    return acc;
  }
  // eq_poly.rs:21-38 (serial in the reference)
  std::vector<Fr> evals() const {
    size_t ell = r.size();
    std::vector<Fr> ev(pow2(ell), Fr::one());
    size_t size = 1;
    for (size_t j = 0; j < ell; j++) {
      size *= 2;
      for (size_t i = size - 1;; i -= 2) {
        Fr scalar = ev[i / 2];
        ev[i] = scalar * r[j];
        ev[i - 1] = scalar - ev[i];
        if (i == 1) break;
      }
    }
    return ev;
  }
  // eq_poly.rs:40-42
  static void compute_factored_lens(size_t ell, size_t& left, size_t& right) {
    left = ell / 2;
    right = ell - ell / 2;
  }
  // eq_poly.rs:44-52
  void compute_factored_evals(std::vector<Fr>& L, std::vector<Fr>& R) const {
    size_t ell = r.size(), left, right;
    compute_factored_lens(ell, left, right);
    L = EqPolynomial(std::vector<Fr>(r.begin(), r.begin() + left)).evals();
    R = EqPolynomial(std::vector<Fr>(r.begin() + left, r.end())).evals();
  }
};

// ---------------------------------------------------------------- poly/commitments.rs
struct MultiCommitGens {
  size_t n = 0;
  std::vector<Point> G;
  Point h;
  // commitments.rs:54-69
  void split_at(size_t mid, MultiCommitGens& a, MultiCommitGens& b) const {
    a.n = mid;
    a.G.assign(G.begin(), G.begin() + mid);
    a.h = h;
    b.n = G.size() - mid;
    b.G.assign(G.begin() + mid, G.end());
    b.h = h;
  }
};

// commitments.rs:22-44 MultiCommitGens::new restated from memory of the absent crates
// (Shake256 -> ChaCha20Rng -> G::rand); yields n+1 prime-order points.  The parity
// contract passes generators explicitly, so only determinism matters here.
inline std::vector<Affine> sample_generators(size_t count, const std::string& label) {
  std::vector<uint8_t> msg(label.begin(), label.end());
  uint8_t gen[32];
  Point::generator().compress(gen);
  msg.insert(msg.end(), gen, gen + 32);
  std::vector<uint8_t> seed = shake256(msg, 32);
  ChaCha20Rng rng(seed.data());
  std::vector<Point> pts;
  for (size_t i = 0; i < count; i++) pts.push_back(point_rand(rng));
  return normalize_batch(pts);
}

inline MultiCommitGens multi_commit_gens_from(const std::vector<Affine>& stream, size_t n) {
  // MultiCommitGens::new(n): samples n+1 points, G = first n, h = the last
  assert(stream.size() >= n + 1);
  MultiCommitGens g;
  g.n = n;
  for (size_t i = 0; i < n; i++) g.G.push_back(Point::from_affine(stream[i]));
  g.h = Point::from_affine(stream[n]);
  return g;
}

// commitments.rs:78-82 Commitments::commit for one scalar
inline Point commit_scalar(const Fr& v, const Fr& blind, const MultiCommitGens& gens) {
  assert(gens.n == 1);
  return gens.G[0] * v + gens.h * blind;
}
// commitments.rs:84-93 batch_commit: re-normalises the generators on every call, exactly
// as the reference does — this is part of the reference's real cost.
inline Point batch_commit(const Fr* inputs, size_t n, const Fr& blind, const MultiCommitGens& gens) {
  assert(gens.n == n);
  std::vector<Affine> bases = normalize_batch(gens.G);
  std::vector<Fr> scalars(inputs, inputs + n);
  bases.push_back(gens.h.into_affine());
  scalars.push_back(blind);
  Point out;
  bool ok = msm(bases, scalars, out);
  assert(ok);
  (void)ok;
  return out;
}

// ---------------------------------------------------------------- subprotocols/dot_product.rs:138-150
struct DotProductProofGens {
  size_t n;
  MultiCommitGens gens_n, gens_1;
  static DotProductProofGens make(size_t n, const std::vector<Affine>& stream) {
    DotProductProofGens g;
    g.n = n;
    multi_commit_gens_from(stream, n + 1).split_at(n, g.gens_n, g.gens_1);
    return g;
  }
};

// ---------------------------------------------------------------- poly/dense_mlpoly.rs
struct PolyCommitmentGens {
  DotProductProofGens gens;
  // dense_mlpoly.rs:40-44
  static PolyCommitmentGens make(size_t num_vars, const std::vector<Affine>& stream) {
    size_t l, r;
    EqPolynomial::compute_factored_lens(num_vars, l, r);
    return PolyCommitmentGens{DotProductProofGens::make(pow2(r), stream)};
  }
};

struct PolyCommitment {
  std::vector<Point> C;
  // dense_mlpoly.rs:281-289
  void append_to_transcript(const char* label, Transcript& t) const {
    t.append_message(label, "poly_commitment_begin");
    for (const Point& p : C) t.append_point("poly_commitment_share", p);
    t.append_message(label, "poly_commitment_end");
  }
};

struct DensePolynomial {
  size_t num_vars = 0, len = 0;
  std::vector<Fr> Z;
  DensePolynomial() {}
  explicit DensePolynomial(std::vector<Fr> z) : Z(std::move(z)) {
    if (Z.empty() || (Z.size() & (Z.size() - 1)))
      throw std::runtime_error("Dense multi-linear polynomials must be made from a power of 2");
    num_vars = log_2(Z.size());
    len = Z.size();
  }
  static DensePolynomial new_padded(std::vector<Fr> ev) {  // dense_mlpoly.rs:74-86
    while (ev.empty() || (ev.size() & (ev.size() - 1))) ev.push_back(Fr::zero());
    return DensePolynomial(ev);
  }
  DensePolynomial clone() const { return DensePolynomial(std::vector<Fr>(Z.begin(), Z.begin() + len)); }
  const Fr& operator[](size_t i) const { return Z[i]; }
  // dense_mlpoly.rs:209-216
  void bound_poly_var_top(const Fr& r) {
    size_t n = len / 2;
    for (size_t i = 0; i < n; i++) Z[i] = Z[i] + r * (Z[i + n] - Z[i]);
    num_vars -= 1;
    len = n;
  }
  // dense_mlpoly.rs:218-225
  void bound_poly_var_bot(const Fr& r) {
    size_t n = len / 2;
    for (size_t i = 0; i < n; i++) Z[i] = Z[2 * i] + r * (Z[2 * i + 1] - Z[2 * i]);
    num_vars -= 1;
    len = n;
  }
  // dense_mlpoly.rs:228-235 + utils/mod.rs:63-73
  Fr evaluate(const std::vector<Fr>& r) const {
    assert(r.size() == num_vars);
    std::vector<Fr> chis = EqPolynomial(r).evals();
    assert(chis.size() == Z.size());
    return compute_dotproduct(Z.data(), chis.data(), Z.size());
  }
  static Fr compute_dotproduct(const Fr* a, const Fr* b, size_t n) {
    Fr total = Fr::zero();
#pragma omp parallel
    {
      Fr local = Fr::zero();
#pragma omp for nowait
      for (size_t i = 0; i < n; i++) local += a[i] * b[i];
#pragma omp critical
      total += local;
    }
    return total;
  }
  // dense_mlpoly.rs:251-261
  static DensePolynomial merge(const std::vector<DensePolynomial>& polys) {
    std::vector<Fr> Z;
    for (const auto& p : polys) Z.insert(Z.end(), p.Z.begin(), p.Z.end());
    Z.resize(next_power_of_two(Z.size()), Fr::zero());
    return DensePolynomial(Z);
  }
  // dense_mlpoly.rs:263-269
  static DensePolynomial from_usize(const std::vector<size_t>& v) {
    std::vector<Fr> Z(v.size());
    for (size_t i = 0; i < v.size(); i++) Z[i] = Fr::from_u64((uint64_t)v[i]);
    return DensePolynomial(Z);
  }
  // dense_mlpoly.rs:152-181 commit (blinds = zeros on this path) + commit_inner 109-128
  PolyCommitment commit(const PolyCommitmentGens& gens) const {
    SpanTimer st("DensePolynomial.commit");
    size_t n = Z.size(), ell = num_vars, lv, rv;
    assert(n == pow2(ell));
    EqPolynomial::compute_factored_lens(ell, lv, rv);
    size_t L_size = pow2(lv), R_size = pow2(rv);
    assert(L_size * R_size == n);
    PolyCommitment pc;
    pc.C.resize(L_size);
    Fr zero = Fr::zero();
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t i = 0; i < L_size; i++)
      pc.C[i] = batch_commit(&Z[R_size * i], R_size, zero, gens.gens.gens_n);
    return pc;
  }
  // dense_mlpoly.rs:183-207
  std::vector<Fr> bound(const std::vector<Fr>& L) const {
    size_t lv, rv;
    EqPolynomial::compute_factored_lens(num_vars, lv, rv);
    size_t L_size = pow2(lv), R_size = pow2(rv);
    std::vector<Fr> out(R_size);
#pragma omp parallel for
    for (size_t i = 0; i < R_size; i++) {
      Fr acc = Fr::zero();
      for (size_t j = 0; j < L_size; j++) acc += L[j] * Z[j * R_size + i];
      out[i] = acc;
    }
    return out;
  }
};

// ---------------------------------------------------------------- utils/gaussian_elimination.rs
inline std::vector<Fr> gaussian_elimination(std::vector<std::vector<Fr>>& m) {
  size_t size = m.size();
  assert(size == m[0].size() - 1);
  auto echelon = [&](size_t i, size_t j) {  // :41-52
    if (m[i][i].is_zero()) return;
    Fr factor = m[j + 1][i] * m[i][i].inverse();
    for (size_t k = i; k < size + 1; k++) {
      Fr tmp = m[i][k];
      m[j + 1][k] -= factor * tmp;
    }
  };
  auto eliminate = [&](size_t i) {  // :54-67
    if (m[i][i].is_zero()) return;
    for (size_t j = i; j >= 1; j--) {
      Fr factor = m[j - 1][i] * m[i][i].inverse();
      for (size_t k = size + 1; k-- > 0;) {
        Fr tmp = m[i][k];
        m[j - 1][k] -= factor * tmp;
      }
    }
  };
  for (size_t i = 0; i + 1 < size; i++)
    for (size_t j = i; j + 1 < size; j++) echelon(i, j);
  for (size_t i = size - 1; i >= 1; i--) eliminate(i);
  std::vector<Fr> result(size);
  for (size_t i = 0; i < size; i++) result[i] = m[i][size] * m[i][i].inverse();
  return result;
}

// ---------------------------------------------------------------- poly/unipoly.rs
struct CompressedUniPoly {
  std::vector<Fr> coeffs_except_linear_term;
};
struct UniPoly {
  std::vector<Fr> coeffs;
  // unipoly.rs:30-54
  static UniPoly from_evals(const std::vector<Fr>& evals) {
    size_t n = evals.size();
    std::vector<std::vector<Fr>> vander;
    for (size_t i = 0; i < n; i++) {
      std::vector<Fr> row;
      Fr x = Fr::from_u64(i);
      row.push_back(Fr::one());
      row.push_back(x);
      for (size_t j = 2; j < n; j++) row.push_back(row[j - 1] * x);
      row.push_back(evals[i]);
      vander.push_back(row);
    }
    return UniPoly{gaussian_elimination(vander)};
  }
  size_t degree() const { return coeffs.size() - 1; }
  Fr eval_at_zero() const { return coeffs[0]; }
  Fr eval_at_one() const {
    Fr s = Fr::zero();
    for (const Fr& c : coeffs) s += c;
    return s;
  }
  // unipoly.rs:72-80
  Fr evaluate(const Fr& r) const {
    Fr eval = coeffs[0], power = r;
    for (size_t i = 1; i < coeffs.size(); i++) {
      eval += power * coeffs[i];
      power *= r;
    }
    return eval;
  }
  // unipoly.rs:82-88
  CompressedUniPoly compress() const {
    CompressedUniPoly c;
    c.coeffs_except_linear_term.push_back(coeffs[0]);
    for (size_t i = 2; i < coeffs.size(); i++) c.coeffs_except_linear_term.push_back(coeffs[i]);
    return c;
  }
  // unipoly.rs:98-109
  static UniPoly decompress(const CompressedUniPoly& c, const Fr& hint) {
    const auto& v = c.coeffs_except_linear_term;
    Fr linear = hint - v[0] - v[0];
    for (size_t i = 1; i < v.size(); i++) linear -= v[i];
    UniPoly p;
    p.coeffs.push_back(v[0]);
    p.coeffs.push_back(linear);
    for (size_t i = 1; i < v.size(); i++) p.coeffs.push_back(v[i]);
    return p;
  }
  // unipoly.rs:112-120
  void append_to_transcript(const char* label, Transcript& t) const {
    t.append_message(label, "UniPoly_begin");
    for (const Fr& c : coeffs) t.append_scalar("coeff", c);
    t.append_message(label, "UniPoly_end");
  }
};

// ---------------------------------------------------------------- subprotocols/sumcheck.rs
struct SumcheckInstanceProof {
  std::vector<CompressedUniPoly> compressed_polys;

  // sumcheck.rs:149-260.  comb_func(vals[alpha]) ; hook (if set) receives each round's evals.
  template <class Func>
  static SumcheckInstanceProof prove_arbitrary(size_t num_rounds, std::vector<DensePolynomial>& polys,
                                               Func comb_func, size_t combined_degree, Transcript& transcript,
                                               std::vector<Fr>& r, std::vector<Fr>& final_evals,
                                               const std::vector<Fr>* scripted_challenges = nullptr,
                                               std::vector<std::vector<Fr>>* round_evals_out = nullptr) {
    SpanTimer st("Sumcheck.prove");
    SumcheckInstanceProof proof;
    size_t alpha = polys.size();
    r.clear();
    for (size_t round = 0; round < num_rounds; round++) {
      std::vector<Fr> eval_points(combined_degree + 1, Fr::zero());
      size_t mle_half = polys[0].len / 2;
#pragma omp parallel
      {
        std::vector<Fr> local(combined_degree + 1, Fr::zero());
        std::vector<Fr> cur(alpha), nxt(alpha);
#pragma omp for nowait
        for (size_t i = 0; i < mle_half; i++) {
          for (size_t j = 0; j < alpha; j++) cur[j] = polys[j][i];
          local[0] += comb_func(cur.data());
          for (size_t j = 0; j < alpha; j++) cur[j] = polys[j][mle_half + i];
          local[1] += comb_func(cur.data());
          for (size_t t = 2; t <= combined_degree; t++) {
            for (size_t j = 0; j < alpha; j++) nxt[j] = cur[j] + polys[j][mle_half + i] - polys[j][i];
            local[t] += comb_func(nxt.data());
            cur.swap(nxt);
          }
        }
#pragma omp critical
        for (size_t t = 0; t <= combined_degree; t++) eval_points[t] += local[t];
      }
      if (round_evals_out) round_evals_out->push_back(eval_points);
      UniPoly round_uni_poly = UniPoly::from_evals(eval_points);
      round_uni_poly.append_to_transcript("poly", transcript);
      Fr r_j = transcript.challenge_scalar("challenge_nextround");
      if (scripted_challenges) r_j = (*scripted_challenges)[round];  // utils/test.rs TestTranscript
      r.push_back(r_j);
      for (auto& p : polys) p.bound_poly_var_top(r_j);  // serial, as sumcheck.rs:251-253
      proof.compressed_polys.push_back(round_uni_poly.compress());
    }
    final_evals.clear();
    for (auto& p : polys) final_evals.push_back(p[0]);
    return proof;
  }

  // sumcheck.rs:26-135; comb_func = A*B*C (grand_product.rs:127-129)
  static SumcheckInstanceProof prove_cubic_batched(const Fr& claim, size_t num_rounds,
                                                   std::vector<DensePolynomial*>& poly_A,
                                                   std::vector<DensePolynomial*>& poly_B,
                                                   DensePolynomial& poly_C, const std::vector<Fr>& coeffs,
                                                   Transcript& transcript, std::vector<Fr>& r,
                                                   std::vector<Fr>& final_A, std::vector<Fr>& final_B,
                                                   Fr& final_C) {
    SpanTimer st("Sumcheck.prove_batched");
    Fr e = claim;
    r.clear();
    SumcheckInstanceProof proof;
    size_t ncirc = poly_A.size();
    for (size_t j = 0; j < num_rounds; j++) {
      std::vector<Fr> e0(ncirc), e2(ncirc), e3(ncirc);
#pragma omp parallel for schedule(dynamic, 1)
      for (size_t k = 0; k < ncirc; k++) {  // rayon across circuits only (sumcheck.rs:50-51)
        const DensePolynomial& A = *poly_A[k];
        const DensePolynomial& B = *poly_B[k];
        Fr p0 = Fr::zero(), p2 = Fr::zero(), p3 = Fr::zero();
        size_t len = A.len / 2;
        for (size_t i = 0; i < len; i++) {
          p0 += A[i] * B[i] * poly_C[i];
          Fr a2 = A[len + i] + A[len + i] - A[i];
          Fr b2 = B[len + i] + B[len + i] - B[i];
          Fr c2 = poly_C[len + i] + poly_C[len + i] - poly_C[i];
          p2 += a2 * b2 * c2;
          Fr a3 = a2 + A[len + i] - A[i];
          Fr b3 = b2 + B[len + i] - B[i];
          Fr c3 = c2 + poly_C[len + i] - poly_C[i];
          p3 += a3 * b3 * c3;
        }
        e0[k] = p0;
        e2[k] = p2;
        e3[k] = p3;
      }
      Fr c0 = Fr::zero(), c2 = Fr::zero(), c3 = Fr::zero();
      for (size_t k = 0; k < ncirc; k++) {
        c0 += e0[k] * coeffs[k];
        c2 += e2[k] * coeffs[k];
        c3 += e3[k] * coeffs[k];
      }
      std::vector<Fr> evals = {c0, e - c0, c2, c3};
      UniPoly poly = UniPoly::from_evals(evals);
      poly.append_to_transcript("poly", transcript);
      Fr r_j = transcript.challenge_scalar("challenge_nextround");
      r.push_back(r_j);
      for (size_t k = 0; k < ncirc; k++) {  // serial binds, sumcheck.rs:116-120
        poly_A[k]->bound_poly_var_top(r_j);
        poly_B[k]->bound_poly_var_top(r_j);
      }
      poly_C.bound_poly_var_top(r_j);
      e = poly.evaluate(r_j);
      proof.compressed_polys.push_back(poly.compress());
    }
    final_A.clear();
    final_B.clear();
    for (size_t k = 0; k < ncirc; k++) {
      final_A.push_back((*poly_A[k])[0]);
      final_B.push_back((*poly_B[k])[0]);
    }
    final_C = poly_C[0];
    return proof;
  }

  // sumcheck.rs:286-328
  bool verify(const Fr& claim, size_t num_rounds, size_t degree_bound, Transcript& transcript, Fr& e_out,
              std::vector<Fr>& r) const {
    Fr e = claim;
    r.clear();
    if (compressed_polys.size() != num_rounds) return false;
    for (size_t i = 0; i < compressed_polys.size(); i++) {
      UniPoly poly = UniPoly::decompress(compressed_polys[i], e);
      if (poly.degree() != degree_bound) return false;
      if (poly.eval_at_zero() + poly.eval_at_one() != e) return false;
      poly.append_to_transcript("poly", transcript);
      Fr r_i = transcript.challenge_scalar("challenge_nextround");
      r.push_back(r_i);
      e = poly.evaluate(r_i);
    }
    e_out = e;
    return true;
  }
};

// ---------------------------------------------------------------- subprotocols/grand_product.rs
struct GrandProductCircuit {
  std::vector<DensePolynomial> left_vec, right_vec;
  // grand_product.rs:38-58 (+ compute_layer 20-36); serial
  explicit GrandProductCircuit(const DensePolynomial& poly) {
    size_t num_layers = log_2(poly.len);
    size_t half = poly.len / 2;
    left_vec.emplace_back(std::vector<Fr>(poly.Z.begin(), poly.Z.begin() + half));
    right_vec.emplace_back(std::vector<Fr>(poly.Z.begin() + half, poly.Z.begin() + 2 * half));
    for (size_t i = 0; i + 1 < num_layers; i++) {
      const DensePolynomial& L = left_vec[i];
      const DensePolynomial& R = right_vec[i];
      size_t len = L.len + R.len;
      std::vector<Fr> ol(len / 4), orr(len / 4);
      for (size_t k = 0; k < len / 4; k++) ol[k] = L[k] * R[k];
      for (size_t k = len / 4; k < len / 2; k++) orr[k - len / 4] = L[k] * R[k];
      left_vec.emplace_back(std::move(ol));
      right_vec.emplace_back(std::move(orr));
    }
  }
  Fr evaluate() const {  // grand_product.rs:60-65
    size_t len = left_vec.size();
    assert(left_vec[len - 1].num_vars == 0);
    return left_vec[len - 1][0] * right_vec[len - 1][0];
  }
};

struct LayerProofBatched {
  SumcheckInstanceProof proof;
  std::vector<Fr> claims_prod_left, claims_prod_right;
};

struct BatchedGrandProductArgument {
  std::vector<LayerProofBatched> proof;

  // grand_product.rs:100-201
  static BatchedGrandProductArgument prove(std::vector<GrandProductCircuit*>& circuits, Transcript& transcript,
                                           std::vector<Fr>& rand_out) {
    SpanTimer st("BatchedGrandProductArgument.prove");
    BatchedGrandProductArgument out;
    size_t num_layers = circuits[0]->left_vec.size();
    std::vector<Fr> claims_to_verify;
    for (auto* c : circuits) claims_to_verify.push_back(c->evaluate());
    std::vector<Fr> rand;
    for (size_t layer_id = num_layers; layer_id-- > 0;) {
      size_t len = circuits[0]->left_vec[layer_id].len + circuits[0]->right_vec[layer_id].len;
      DensePolynomial poly_C(EqPolynomial(rand).evals());
      assert(poly_C.len == len / 2);
      (void)len;
      size_t num_rounds_prod = log_2(poly_C.len);
      std::vector<DensePolynomial*> A, B;
      for (auto* c : circuits) {
        A.push_back(&c->left_vec[layer_id]);
        B.push_back(&c->right_vec[layer_id]);
      }
      std::vector<Fr> coeff_vec = transcript.challenge_vector("rand_coeffs_next_layer", claims_to_verify.size());
      Fr claim = Fr::zero();
      for (size_t i = 0; i < claims_to_verify.size(); i++) claim += claims_to_verify[i] * coeff_vec[i];
      std::vector<Fr> rand_prod, cl, cr;
      Fr ceq;
      LayerProofBatched lp;
      lp.proof = SumcheckInstanceProof::prove_cubic_batched(claim, num_rounds_prod, A, B, poly_C, coeff_vec,
                                                            transcript, rand_prod, cl, cr, ceq);
      for (size_t i = 0; i < circuits.size(); i++) {
        transcript.append_scalar("claim_prod_left", cl[i]);
        transcript.append_scalar("claim_prod_right", cr[i]);
      }
      Fr r_layer = transcript.challenge_scalar("challenge_r_layer");
      claims_to_verify.clear();
      for (size_t i = 0; i < circuits.size(); i++) claims_to_verify.push_back(cl[i] + r_layer * (cr[i] - cl[i]));
      std::vector<Fr> ext = {r_layer};
      ext.insert(ext.end(), rand_prod.begin(), rand_prod.end());
      rand = ext;
      lp.claims_prod_left = cl;
      lp.claims_prod_right = cr;
      out.proof.push_back(std::move(lp));
    }
    rand_out = rand;
    return out;
  }

  // grand_product.rs:203-261
  bool verify(const std::vector<Fr>& claims_prod_vec, size_t len, Transcript& transcript,
              std::vector<Fr>& claims_out, std::vector<Fr>& rand_out) const {
    size_t num_layers = log_2(len);
    std::vector<Fr> rand;
    if (proof.size() != num_layers) return false;
    std::vector<Fr> claims_to_verify = claims_prod_vec;
    for (size_t i = 0; i < num_layers; i++) {
      size_t num_rounds = i;
      std::vector<Fr> coeff_vec = transcript.challenge_vector("rand_coeffs_next_layer", claims_to_verify.size());
      Fr claim = Fr::zero();
      for (size_t k = 0; k < claims_to_verify.size(); k++) claim += claims_to_verify[k] * coeff_vec[k];
      Fr claim_last;
      std::vector<Fr> rand_prod;
      if (!proof[i].proof.verify(claim, num_rounds, 3, transcript, claim_last, rand_prod)) return false;
      const auto& cl = proof[i].claims_prod_left;
      const auto& cr = proof[i].claims_prod_right;
      if (cl.size() != claims_prod_vec.size() || cr.size() != claims_prod_vec.size()) return false;
      for (size_t k = 0; k < claims_prod_vec.size(); k++) {
        transcript.append_scalar("claim_prod_left", cl[k]);
        transcript.append_scalar("claim_prod_right", cr[k]);
      }
      if (rand.size() != rand_prod.size()) return false;
      Fr eq = Fr::one();
      for (size_t k = 0; k < rand.size(); k++)
        eq = eq * (rand[k] * rand_prod[k] + (Fr::one() - rand[k]) * (Fr::one() - rand_prod[k]));
      Fr claim_expected = Fr::zero();
      for (size_t k = 0; k < claims_prod_vec.size(); k++) claim_expected += coeff_vec[k] * (cl[k] * cr[k] * eq);
      if (claim_expected != claim_last) return false;
      Fr r_layer = transcript.challenge_scalar("challenge_r_layer");
      claims_to_verify.clear();
      for (size_t k = 0; k < cl.size(); k++) claims_to_verify.push_back(cl[k] + r_layer * (cr[k] - cl[k]));
      std::vector<Fr> ext = {r_layer};
      ext.insert(ext.end(), rand_prod.begin(), rand_prod.end());
      rand = ext;
    }
    claims_out = claims_to_verify;
    rand_out = rand;
    return true;
  }
};

// ---------------------------------------------------------------- subprotocols/bullet.rs
inline Fr inner_product(const Fr* a, const Fr* b, size_t n) {  // bullet.rs:265-275
  Fr out = Fr::zero();
  for (size_t i = 0; i < n; i++) out += a[i] * b[i];
  return out;
}

struct BulletReductionProof {
  std::vector<Point> L_vec, R_vec;

  // bullet.rs:40-154 (serial, as the reference)
  static BulletReductionProof prove(Transcript& transcript, const Point& Q, const std::vector<Point>& G_vec,
                                    const Point& H, const std::vector<Fr>& a_vec, const std::vector<Fr>& b_vec,
                                    const Fr& blind, const std::vector<std::pair<Fr, Fr>>& blinds_vec,
                                    Point& Gamma_hat, Fr& a_hat, Fr& b_hat, Point& g_hat, Fr& blind_fin_out) {
    std::vector<Point> G = G_vec;
    std::vector<Fr> a = a_vec, b = b_vec;
    size_t n = G.size();
    assert((n & (n - 1)) == 0);
    size_t lg_n = log_2(n);
    assert(a.size() == n && b.size() == n && blinds_vec.size() == 2 * lg_n);
    (void)lg_n;
    BulletReductionProof proof;
    size_t bi = 0;
    Fr blind_fin = blind;
    while (n != 1) {
      n /= 2;
      Fr c_L = inner_product(&a[0], &b[n], n);
      Fr c_R = inner_product(&a[n], &b[0], n);
      const Fr& blind_L = blinds_vec[bi].first;
      const Fr& blind_R = blinds_vec[bi].second;
      bi++;
      std::vector<Fr> scalars(a.begin(), a.begin() + n);
      scalars.push_back(c_L);
      scalars.push_back(blind_L);
      std::vector<Point> bases(G.begin() + n, G.begin() + 2 * n);
      bases.push_back(Q);
      bases.push_back(H);
      Point L, R;
      msm(normalize_batch(bases), scalars, L);
      scalars.assign(a.begin() + n, a.begin() + 2 * n);
      scalars.push_back(c_R);
      scalars.push_back(blind_R);
      bases.assign(G.begin(), G.begin() + n);
      bases.push_back(Q);
      bases.push_back(H);
      msm(normalize_batch(bases), scalars, R);
      transcript.append_point("L", L);
      transcript.append_point("R", R);
      Fr u = transcript.challenge_scalar("u");
      Fr u_inv = u.inverse();
      for (size_t i = 0; i < n; i++) {
        a[i] = a[i] * u + u_inv * a[n + i];
        b[i] = b[i] * u_inv + u * b[n + i];
        G[i] = G[i] * u_inv + G[n + i] * u;
      }
      blind_fin = blind_fin + blind_L * u * u + blind_R * u_inv * u_inv;
      proof.L_vec.push_back(L);
      proof.R_vec.push_back(R);
    }
    Gamma_hat = G[0] * a[0] + Q * (a[0] * b[0]) + H * blind_fin;
    a_hat = a[0];
    b_hat = b[0];
    g_hat = G[0];
    blind_fin_out = blind_fin;
    return proof;
  }

  // bullet.rs:159-221
  bool verification_scalars(size_t n, Transcript& transcript, std::vector<Fr>& u_sq, std::vector<Fr>& u_inv_sq,
                            std::vector<Fr>& s) const {
    size_t lg_n = L_vec.size();
    if (lg_n >= 32) return false;
    if (n != ((size_t)1 << lg_n)) return false;
    std::vector<Fr> challenges;
    for (size_t i = 0; i < lg_n; i++) {
      transcript.append_point("L", L_vec[i]);
      transcript.append_point("R", R_vec[i]);
      challenges.push_back(transcript.challenge_scalar("u"));
    }
    std::vector<Fr> challenges_inv;
    for (auto& c : challenges) challenges_inv.push_back(c.inverse());
    Fr all_inv = Fr::one();
    for (auto& c : challenges_inv) all_inv *= c;
    for (size_t i = 0; i < lg_n; i++) {
      challenges[i] = challenges[i].square();
      challenges_inv[i] = challenges_inv[i].square();
    }
    u_sq = challenges;
    u_inv_sq = challenges_inv;
    s.clear();
    s.push_back(all_inv);
    for (size_t i = 1; i < n; i++) {
      size_t lg_i = 31 - __builtin_clz((uint32_t)i);
      size_t k = (size_t)1 << lg_i;
      s.push_back(s[i - k] * u_sq[(lg_n - 1) - lg_i]);
    }
    return true;
  }
  // bullet.rs:227-257
  bool verify(size_t n, const std::vector<Fr>& a, Transcript& transcript, const Point& Gamma,
              const std::vector<Point>& G, Point& G_hat, Point& Gamma_hat, Fr& a_hat) const {
    std::vector<Fr> u_sq, u_inv_sq, s;
    if (!verification_scalars(n, transcript, u_sq, u_inv_sq, s)) return false;
    msm(normalize_batch(G), s, G_hat);
    a_hat = inner_product(a.data(), s.data(), n);
    std::vector<Point> bases = L_vec;
    bases.insert(bases.end(), R_vec.begin(), R_vec.end());
    bases.push_back(Gamma);
    std::vector<Fr> scalars = u_sq;
    scalars.insert(scalars.end(), u_inv_sq.begin(), u_inv_sq.end());
    scalars.push_back(Fr::one());
    msm(normalize_batch(bases), scalars, Gamma_hat);
    return true;
  }
};

// ---------------------------------------------------------------- subprotocols/dot_product.rs:152-297
struct DotProductProofLog {
  BulletReductionProof bullet_reduction_proof;
  Point delta, beta;
  Fr z1, z2;

  static DotProductProofLog prove(const DotProductProofGens& gens, Transcript& transcript, RandomTape& tape,
                                  const std::vector<Fr>& x_vec, const Fr& blind_x, const std::vector<Fr>& a_vec,
                                  const Fr& y, const Fr& blind_y, Point& Cx, Point& Cy) {
    SpanTimer st("DotProductProofLog.prove");
    transcript.append_protocol_name("dot product proof (log)");
    size_t n = x_vec.size();
    assert(a_vec.size() == n && gens.n == n);
    Fr d = tape.random_scalar("d");
    Fr r_delta = tape.random_scalar("r_delta");
    Fr r_beta = tape.random_scalar("r_delta");  // sic: dot_product.rs:189 reuses the label
    std::vector<Fr> v1 = tape.random_vector("blinds_vec_1", 2 * log_2(n));
    std::vector<Fr> v2 = tape.random_vector("blinds_vec_2", 2 * log_2(n));
    std::vector<std::pair<Fr, Fr>> blinds_vec;
    for (size_t i = 0; i < v1.size(); i++) blinds_vec.push_back({v1[i], v2[i]});
    Cx = batch_commit(x_vec.data(), n, blind_x, gens.gens_n);
    transcript.append_point("Cx", Cx);
    Cy = commit_scalar(y, blind_y, gens.gens_1);
    transcript.append_point("Cy", Cy);
    transcript.append_scalars("a", a_vec);
    Fr blind_Gamma = blind_x + blind_y;
    DotProductProofLog out;
    Point Gamma_hat, g_hat;
    Fr x_hat, a_hat, rhat_Gamma;
    out.bullet_reduction_proof =
        BulletReductionProof::prove(transcript, gens.gens_1.G[0], gens.gens_n.G, gens.gens_n.h, x_vec, a_vec,
                                    blind_Gamma, blinds_vec, Gamma_hat, x_hat, a_hat, g_hat, rhat_Gamma);
    Fr y_hat = x_hat * a_hat;
    MultiCommitGens gens_hat;
    gens_hat.n = 1;
    gens_hat.G = {g_hat};
    gens_hat.h = gens.gens_1.h;
    out.delta = commit_scalar(d, r_delta, gens_hat);
    transcript.append_point("delta", out.delta);
    out.beta = commit_scalar(d, r_beta, gens.gens_1);
    transcript.append_point("beta", out.beta);
    Fr c = transcript.challenge_scalar("c");
    out.z1 = d + c * y_hat;
    out.z2 = a_hat * (c * rhat_Gamma + r_beta) + r_delta;
    return out;
  }

  bool verify(size_t n, const DotProductProofGens& gens, Transcript& transcript, const std::vector<Fr>& a,
              const Point& Cx, const Point& Cy) const {
    if (gens.n != n || a.size() != n) return false;
    transcript.append_protocol_name("dot product proof (log)");
    transcript.append_point("Cx", Cx);
    transcript.append_point("Cy", Cy);
    transcript.append_scalars("a", a);
    Point Gamma = Cx + Cy;
    Point g_hat, Gamma_hat;
    Fr a_hat;
    if (!bullet_reduction_proof.verify(n, a, transcript, Gamma, gens.gens_n.G, g_hat, Gamma_hat, a_hat)) return false;
    transcript.append_point("delta", delta);
    transcript.append_point("beta", beta);
    Fr c = transcript.challenge_scalar("c");
    Point lhs = (Gamma_hat * c + beta) * a_hat + delta;
    Point rhs = (g_hat + gens.gens_1.G[0] * a_hat) * z1 + gens.gens_1.h * z2;
    return lhs == rhs;
  }
};

// ---------------------------------------------------------------- dense_mlpoly.rs:291-401 PolyEvalProof
struct PolyEvalProof {
  DotProductProofLog proof;

  static PolyEvalProof prove(const DensePolynomial& poly, const std::vector<Fr>& r, const Fr& Zr,
                             const PolyCommitmentGens& gens, Transcript& transcript, RandomTape& tape) {
    SpanTimer st("DensePolyEval.prove");
    transcript.append_protocol_name("polynomial evaluation proof");
    assert(poly.num_vars == r.size());
    std::vector<Fr> L, R;
    EqPolynomial(r).compute_factored_evals(L, R);
    std::vector<Fr> LZ = poly.bound(L);
    Fr LZ_blind = Fr::zero();  // blinds are all zero on this path (dense_mlpoly.rs:325-344)
    Fr blind_Zr = Fr::zero();
    Point Cx, Cy;
    PolyEvalProof out;
    out.proof = DotProductProofLog::prove(gens.gens, transcript, tape, LZ, LZ_blind, R, Zr, blind_Zr, Cx, Cy);
    return out;
  }
  // dense_mlpoly.rs:361-386
  bool verify(const PolyCommitmentGens& gens, Transcript& transcript, const std::vector<Fr>& r, const Point& C_Zr,
              const PolyCommitment& comm) const {
    transcript.append_protocol_name("polynomial evaluation proof");
    std::vector<Fr> L, R;
    EqPolynomial(r).compute_factored_evals(L, R);
    Point C_LZ;
    if (!msm(normalize_batch(comm.C), L, C_LZ)) return false;
    return proof.verify(R.size(), gens.gens, transcript, R, C_LZ, C_Zr);
  }
  // dense_mlpoly.rs:388-400
  bool verify_plain(const PolyCommitmentGens& gens, Transcript& transcript, const std::vector<Fr>& r, const Fr& Zr,
                    const PolyCommitment& comm) const {
    Point C_Zr = commit_scalar(Zr, Fr::zero(), gens.gens.gens_1);
    return verify(gens, transcript, r, C_Zr, comm);
  }
};

// ---------------------------------------------------------------- subtables/*.rs
enum StrategyKind { STRAT_AND = 0, STRAT_OR = 1, STRAT_XOR = 2, STRAT_LT = 3, STRAT_RANGE = 4 };

// Runtime stand-in for `impl SubtableStrategy<F, C, M> for ...` (const generics in the reference).
struct Strategy {
  int kind;
  size_t C, log_m, log_r;  // M = 2^log_m; log_r only for RangeCheckSubtableStrategy<LOG_R>
  size_t M() const { return pow2(log_m); }
  size_t num_subtables() const { return kind == STRAT_LT ? 2 : (kind == STRAT_RANGE ? 3 : 1); }
  size_t num_memories() const { return kind == STRAT_LT ? 2 * C : C; }
  size_t g_poly_degree() const { return kind == STRAT_LT ? C : 1; }
  size_t sumcheck_poly_degree() const { return g_poly_degree() + 1; }
  // subtables/mod.rs:64-74, range_check.rs:62-73
  size_t memory_to_subtable_index(size_t i) const {
    if (kind == STRAT_RANGE) {
      if (i * log_m > log_r) return 2;
      return ((i + 1) * log_m > log_r) ? 1 : 0;
    }
    return i % num_subtables();
  }
  size_t memory_to_dimension_index(size_t i) const {
    if (kind == STRAT_RANGE) return i;
    return i / num_subtables();
  }
  // and.rs:16-28, or.rs, xor.rs:16-27, lt.rs:16-30, range_check.rs:15-34
  std::vector<std::vector<Fr>> materialize_subtables() const {
    size_t m = M();
    std::vector<std::vector<Fr>> out;
    if (kind == STRAT_RANGE) {
      std::vector<Fr> full(m), rem(m), zeros(m, Fr::zero());
      size_t cutoff = (size_t)1 << (log_r % log_m);
      for (size_t i = 0; i < m; i++) {
        full[i] = Fr::from_u64(i);
        rem[i] = i < cutoff ? Fr::from_u64(i) : Fr::zero();
      }
      return {full, rem, zeros};
    }
    size_t bits = log_m / 2;
    if (kind == STRAT_LT) {
      std::vector<Fr> lt(m), eq(m);
      for (size_t idx = 0; idx < m; idx++) {
        size_t lhs, rhs;
        split_bits(idx, bits, lhs, rhs);
        lt[idx] = Fr::from_u64(lhs < rhs);
        eq[idx] = Fr::from_u64(lhs == rhs);
      }
      return {lt, eq};
    }
    std::vector<Fr> t(m);
    for (size_t idx = 0; idx < m; idx++) {
      size_t lhs, rhs;
      split_bits(idx, bits, lhs, rhs);
      size_t v = kind == STRAT_AND ? (lhs & rhs) : (kind == STRAT_OR ? (lhs | rhs) : (lhs ^ rhs));
      t[idx] = Fr::from_u64(v);
    }
    return {t};
  }
  // evaluate_subtable_mle: and.rs:30-40, or.rs, xor.rs:29-42, lt.rs:33-55, range_check.rs:36-60
  Fr evaluate_subtable_mle(size_t subtable_index, const std::vector<Fr>& point) const {
    Fr one = Fr::one();
    if (kind == STRAT_RANGE) {
      size_t b = point.size();
      if (subtable_index == 0) {
        Fr res = Fr::zero();
        for (size_t i = 0; i < b; i++) res += Fr::from_u64(1ULL << i) * point[b - i - 1];
        return res;
      } else if (subtable_index == 1) {
        size_t cutoff = log_r % log_m;
        Fr res = Fr::zero();
        for (size_t i = 0; i < b; i++) {
          if (i < cutoff)
            res += Fr::from_u64(1ULL << i) * point[b - i - 1];
          else
            res *= one - point[b - i - 1];
        }
        return res;
      }
      return Fr::zero();
    }
    size_t b = point.size() / 2;
    const Fr* x = &point[0];
    const Fr* y = &point[b];
    if (kind == STRAT_LT) {
      Fr eq_term = one;
      if (subtable_index % 2 == 0) {
        Fr res = Fr::zero();
        for (size_t i = 0; i < b; i++) {
          res += (one - x[i]) * y[i] * eq_term;
          eq_term *= one - x[i] - y[i] + Fr::from_u64(2) * x[i] * y[i];
        }
        return res;
      }
      for (size_t i = 0; i < b; i++) eq_term *= one - x[i] - y[i] + Fr::from_u64(2) * x[i] * y[i];
      return eq_term;
    }
    Fr res = Fr::zero();
    for (size_t i = 0; i < b; i++) {
      Fr xv = x[b - i - 1], yv = y[b - i - 1];
      Fr term;
      if (kind == STRAT_AND)
        term = xv * yv;
      else if (kind == STRAT_OR)
        term = one - (one - xv) * (one - yv);
      else
        term = (one - xv) * yv + xv * (one - yv);
      res += Fr::from_u64(1ULL << i) * term;
    }
    return res;
  }
  // combine_lookups: and.rs:45-53, lt.rs:60-69, range_check.rs:78-86
  Fr combine_lookups(const Fr* vals) const {
    if (kind == STRAT_LT) {
      Fr sum = Fr::zero(), eq_prod = Fr::one();
      for (size_t i = 0; i < C; i++) {
        sum += vals[2 * i] * eq_prod;
        eq_prod *= vals[2 * i + 1];
      }
      return sum;
    }
    size_t increment = kind == STRAT_RANGE ? log_m : log_m / 2;
    Fr sum = Fr::zero();
    for (size_t i = 0; i < num_memories(); i++) sum += Fr::from_u64(1ULL << (i * increment)) * vals[i];
    return sum;
  }
  // subtables/mod.rs:53-57
  Fr combine_lookups_eq(const Fr* vals) const { return combine_lookups(vals) * vals[num_memories()]; }
};

// ---------------------------------------------------------------- lasso/densified.rs
struct DensifiedRepresentation {
  std::vector<std::vector<size_t>> dim_usize;
  std::vector<DensePolynomial> dim, read, final_;
  DensePolynomial combined_l_variate_polys, combined_log_m_variate_polys;
  size_t s, log_m, m, C;

  // densified.rs:21-75; indices[j][i] = lookup j, dimension i
  static DensifiedRepresentation from_lookup_indices(const std::vector<std::vector<size_t>>& indices, size_t C,
                                                     size_t log_m) {
    SpanTimer st("Densify");
    DensifiedRepresentation d;
    d.C = C;
    d.s = next_power_of_two(indices.size());
    d.log_m = log_m;
    d.m = pow2(log_m);
    for (size_t i = 0; i < C; i++) {
      std::vector<size_t> access_sequence(indices.size());
      for (size_t j = 0; j < indices.size(); j++) access_sequence[j] = indices[j][i];
      access_sequence.resize(d.s, 0);
      std::vector<size_t> final_timestamps(d.m, 0), read_timestamps(d.s, 0);
      for (size_t k = 0; k < d.s; k++) {
        size_t addr = access_sequence[k];
        assert(addr < d.m);
        size_t ts = final_timestamps[addr];
        read_timestamps[k] = ts;
        final_timestamps[addr] = ts + 1;
      }
      d.dim.push_back(DensePolynomial::from_usize(access_sequence));
      d.read.push_back(DensePolynomial::from_usize(read_timestamps));
      d.final_.push_back(DensePolynomial::from_usize(final_timestamps));
      d.dim_usize.push_back(access_sequence);
    }
    std::vector<DensePolynomial> l_variate = d.dim;
    l_variate.insert(l_variate.end(), d.read.begin(), d.read.end());
    d.combined_l_variate_polys = DensePolynomial::merge(l_variate);
    d.combined_log_m_variate_polys = DensePolynomial::merge(d.final_);
    return d;
  }
};

// ---------------------------------------------------------------- lasso/surge.rs:25-82
struct SparsePolyCommitmentGens {
  PolyCommitmentGens gens_combined_l_variate, gens_combined_log_m_variate, gens_derefs;
  // surge.rs:32-58.  `stream` = the label's generator stream (explicit input);
  // needs_points() tells the caller how many are required.
  static void num_vars(size_t c, size_t s, size_t num_memories, size_t log_m, size_t& nv_l, size_t& nv_m,
                       size_t& nv_d) {
    nv_l = log_2(next_power_of_two(2 * c * s));
    nv_m = log_2(next_power_of_two(c)) + log_m;
    nv_d = log_2(next_power_of_two(num_memories * s));
  }
  static size_t needs_points(size_t c, size_t s, size_t num_memories, size_t log_m) {
    size_t a, b, d;
    num_vars(c, s, num_memories, log_m, a, b, d);
    size_t mx = std::max(a, std::max(b, d));
    return pow2(mx - mx / 2) + 2;
  }
  static SparsePolyCommitmentGens make(size_t c, size_t s, size_t num_memories, size_t log_m,
                                       const std::vector<Affine>& stream) {
    size_t a, b, d;
    num_vars(c, s, num_memories, log_m, a, b, d);
    return SparsePolyCommitmentGens{PolyCommitmentGens::make(a, stream), PolyCommitmentGens::make(b, stream),
                                    PolyCommitmentGens::make(d, stream)};
  }
};

struct SparsePolynomialCommitment {
  PolyCommitment l_variate_polys_commitment, log_m_variate_polys_commitment;
  size_t s, log_m, m;
};

// densified.rs:77-96
inline SparsePolynomialCommitment densified_commit(const DensifiedRepresentation& d,
                                                   const SparsePolyCommitmentGens& gens) {
  SpanTimer st("DensifiedRepresentation.commit");
  SparsePolynomialCommitment c;
  c.l_variate_polys_commitment = d.combined_l_variate_polys.commit(gens.gens_combined_l_variate);
  c.log_m_variate_polys_commitment = d.combined_log_m_variate_polys.commit(gens.gens_combined_log_m_variate);
  c.s = d.s;
  c.log_m = d.log_m;
  c.m = d.m;
  return c;
}

// ---------------------------------------------------------------- lasso/memory_checking.rs:149-310
struct GrandProducts {
  GrandProductCircuit init, read, write, final_;
};

inline GrandProducts make_grand_products(const std::vector<Fr>& eval_table, const DensePolynomial& dim_i,
                                         const std::vector<size_t>& dim_i_usize, const DensePolynomial& read_i,
                                         const DensePolynomial& final_i, const Fr& gamma, const Fr& tau) {
  // hash(a, v, t) = t * gamma^2 + v * gamma + a - tau   (memory_checking.rs:249-252)
  Fr g2 = gamma.square();
  auto hash_func = [&](const Fr& a, const Fr& v, const Fr& t) { return t * g2 + v * gamma + a - tau; };
  assert(eval_table.size() == final_i.len);
  size_t num_mem_cells = eval_table.size();
  std::vector<Fr> init(num_mem_cells), fin(num_mem_cells);
  for (size_t i = 0; i < num_mem_cells; i++) {
    init[i] = hash_func(Fr::from_u64(i), eval_table[i], Fr::zero());
    fin[i] = hash_func(Fr::from_u64(i), eval_table[i], final_i[i]);
  }
  assert(dim_i.len == read_i.len);
  size_t num_ops = dim_i.len;
  std::vector<Fr> rd(num_ops), wr(num_ops);
#pragma omp parallel for
  for (size_t i = 0; i < num_ops; i++) {
    rd[i] = hash_func(dim_i[i], eval_table[dim_i_usize[i]], read_i[i]);
    wr[i] = hash_func(dim_i[i], eval_table[dim_i_usize[i]], read_i[i] + Fr::one());
  }
  return GrandProducts{GrandProductCircuit(DensePolynomial(init)), GrandProductCircuit(DensePolynomial(rd)),
                       GrandProductCircuit(DensePolynomial(wr)), GrandProductCircuit(DensePolynomial(fin))};
}

// ---------------------------------------------------------------- subtables/mod.rs:95-216
struct Subtables {
  Strategy S;
  std::vector<std::vector<Fr>> subtable_entries;
  std::vector<DensePolynomial> lookup_polys;
  DensePolynomial combined_poly;

  // subtables/mod.rs:116-129 + to_lookup_polys 78-92
  Subtables(const Strategy& strat, const std::vector<std::vector<size_t>>& nz, size_t s) : S(strat) {
    for (auto& d : nz) {
      assert(d.size() == s);
      (void)d;
    }
    subtable_entries = S.materialize_subtables();
    for (size_t i = 0; i < S.num_memories(); i++) {
      std::vector<Fr> lookups(s);
      const auto& subtable = subtable_entries[S.memory_to_subtable_index(i)];
      const auto& idx = nz[S.memory_to_dimension_index(i)];
      for (size_t j = 0; j < s; j++) lookups[j] = subtable[idx[j]];
      lookup_polys.emplace_back(std::move(lookups));
    }
    combined_poly = DensePolynomial::merge(lookup_polys);
  }
  // subtables/mod.rs:133-175 (rayon over memories)
  std::vector<GrandProducts> to_grand_products(const DensifiedRepresentation& dense, const Fr& gamma,
                                               const Fr& tau) const {
    SpanTimer st("Subtables.to_grand_products");
    size_t nm = S.num_memories();
    std::vector<GrandProducts*> tmp(nm, nullptr);
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t i = 0; i < nm; i++) {
      size_t j = S.memory_to_dimension_index(i);
      tmp[i] = new GrandProducts(make_grand_products(subtable_entries[S.memory_to_subtable_index(i)], dense.dim[j],
                                                     dense.dim_usize[j], dense.read[j], dense.final_[j], gamma, tau));
    }
    std::vector<GrandProducts> out;
    for (size_t i = 0; i < nm; i++) {
      out.push_back(std::move(*tmp[i]));
      delete tmp[i];
    }
    return out;
  }
  // subtables/mod.rs:177-184
  PolyCommitment commit(const PolyCommitmentGens& gens) const {
    SpanTimer st("Subtables.commit");
    return combined_poly.commit(gens);
  }
  // subtables/mod.rs:186-216
  Fr compute_sumcheck_claim(const EqPolynomial& eq) const {
    SpanTimer st("Subtables.compute_sumcheck_claim");
    size_t hypercube_size = lookup_polys[0].len;
    std::vector<Fr> eq_evals = eq.evals();
    size_t nm = S.num_memories();
    Fr total = Fr::zero();
#pragma omp parallel
    {
      Fr local = Fr::zero();
      std::vector<Fr> ops(nm);
#pragma omp for nowait
      for (size_t k = 0; k < hypercube_size; k++) {
        for (size_t j = 0; j < nm; j++) ops[j] = lookup_polys[j][k];
        local += eq_evals[k] * S.combine_lookups(ops.data());
      }
#pragma omp critical
      total += local;
    }
    return total;
  }
};

// subtables/mod.rs:225-380 CombinedTableEvalProof
struct CombinedTableEvalProof {
  PolyEvalProof proof_table_eval;

  static CombinedTableEvalProof prove(const DensePolynomial& combined_poly, const std::vector<Fr>& eval_ops_val_vec,
                                      const std::vector<Fr>& r, const PolyCommitmentGens& gens,
                                      Transcript& transcript, RandomTape& tape) {
    SpanTimer st("CombinedEval.prove");
    transcript.append_protocol_name("Lasso CombinedTableEvalProof");
    std::vector<Fr> evals = eval_ops_val_vec;
    evals.resize(next_power_of_two(evals.size()), Fr::zero());
    // prove_single (mod.rs:230-281)
    assert(combined_poly.num_vars == r.size() + log_2(evals.size()));
    transcript.append_scalars("evals_ops_val", evals);
    std::vector<Fr> challenges = transcript.challenge_vector("challenge_combine_n_to_one", log_2(evals.size()));
    DensePolynomial poly_evals(evals);
    for (size_t i = challenges.size(); i-- > 0;) poly_evals.bound_poly_var_bot(challenges[i]);
    assert(poly_evals.len == 1);
    Fr joint_claim_eval = poly_evals[0];
    std::vector<Fr> r_joint = challenges;
    r_joint.insert(r_joint.end(), r.begin(), r.end());
    transcript.append_scalar("joint_claim_eval", joint_claim_eval);
    CombinedTableEvalProof out;
    out.proof_table_eval = PolyEvalProof::prove(combined_poly, r_joint, joint_claim_eval, gens, transcript, tape);
    return out;
  }
  // mod.rs:315-375
  bool verify(const std::vector<Fr>& r, const std::vector<Fr>& evals_in, const PolyCommitmentGens& gens,
              const PolyCommitment& comm, Transcript& transcript) const {
    transcript.append_protocol_name("Lasso CombinedTableEvalProof");
    std::vector<Fr> evals = evals_in;
    evals.resize(next_power_of_two(evals.size()), Fr::zero());
    transcript.append_scalars("evals_ops_val", evals);
    std::vector<Fr> challenges = transcript.challenge_vector("challenge_combine_n_to_one", log_2(evals.size()));
    DensePolynomial poly_evals(evals);
    for (size_t i = challenges.size(); i-- > 0;) poly_evals.bound_poly_var_bot(challenges[i]);
    Fr joint_claim_eval = poly_evals[0];
    std::vector<Fr> r_joint = challenges;
    r_joint.insert(r_joint.end(), r.begin(), r.end());
    transcript.append_scalar("joint_claim_eval", joint_claim_eval);
    return proof_table_eval.verify_plain(gens, transcript, r_joint, joint_claim_eval, comm);
  }
};

// mod.rs:382-393 CombinedTableCommitment::append_to_transcript
inline void append_combined_table_commitment(const PolyCommitment& comm_ops_val, const char* label, Transcript& t) {
  t.append_message("subtable_evals_commitment", "begin_subtable_evals_commitment");
  comm_ops_val.append_to_transcript(label, t);
  t.append_message("subtable_evals_commitment", "end_subtable_evals_commitment");
}

// ---------------------------------------------------------------- memory_checking.rs:655-785 ProductLayerProof
struct ProductLayerProof {
  std::vector<std::array<Fr, 4>> grand_product_evals;  // (init, read, write, final) per memory
  BatchedGrandProductArgument proof_mem, proof_ops;

  static ProductLayerProof prove(std::vector<GrandProducts>& gps, Transcript& transcript, std::vector<Fr>& rand_mem,
                                 std::vector<Fr>& rand_ops) {
    SpanTimer st("ProductLayer.prove");
    transcript.append_protocol_name("Lasso ProductLayerProof");
    ProductLayerProof out;
    for (auto& gp : gps) {
      Fr hi = gp.init.evaluate(), hr = gp.read.evaluate(), hw = gp.write.evaluate(), hf = gp.final_.evaluate();
      if (hi * hw != hr * hf) throw std::runtime_error("multiset hash mismatch (memory_checking.rs:689)");
      transcript.append_scalar("claim_hash_init", hi);
      transcript.append_scalar("claim_hash_read", hr);
      transcript.append_scalar("claim_hash_write", hw);
      transcript.append_scalar("claim_hash_final", hf);
      out.grand_product_evals.push_back({hi, hr, hw, hf});
    }
    std::vector<GrandProductCircuit*> rw, inf;
    for (auto& gp : gps) {
      rw.push_back(&gp.read);
      rw.push_back(&gp.write);
    }
    out.proof_ops = BatchedGrandProductArgument::prove(rw, transcript, rand_ops);
    for (auto& gp : gps) {
      inf.push_back(&gp.init);
      inf.push_back(&gp.final_);
    }
    out.proof_mem = BatchedGrandProductArgument::prove(inf, transcript, rand_mem);
    return out;
  }
  bool verify(size_t num_ops, size_t num_cells, Transcript& transcript, std::vector<Fr>& claims_mem,
              std::vector<Fr>& rand_mem, std::vector<Fr>& claims_ops, std::vector<Fr>& rand_ops) const {
    transcript.append_protocol_name("Lasso ProductLayerProof");
    std::vector<Fr> rw_claims, if_claims;
    for (auto& e : grand_product_evals) {
      if (e[0] * e[2] != e[1] * e[3]) return false;
      transcript.append_scalar("claim_hash_init", e[0]);
      transcript.append_scalar("claim_hash_read", e[1]);
      transcript.append_scalar("claim_hash_write", e[2]);
      transcript.append_scalar("claim_hash_final", e[3]);
      rw_claims.push_back(e[1]);
      rw_claims.push_back(e[2]);
      if_claims.push_back(e[0]);
      if_claims.push_back(e[3]);
    }
    if (!proof_ops.verify(rw_claims, num_ops, transcript, claims_ops, rand_ops)) return false;
    if (!proof_mem.verify(if_claims, num_cells, transcript, claims_mem, rand_mem)) return false;
    return true;
  }
};

// identity_poly.rs:14-20
inline Fr identity_poly_evaluate(const std::vector<Fr>& r) {
  size_t len = r.size();
  Fr s = Fr::zero();
  for (size_t i = 0; i < len; i++) s += Fr::from_u64((uint64_t)pow2(len - i - 1)) * r[i];
  return s;
}

// ---------------------------------------------------------------- memory_checking.rs:313-653 HashLayerProof
struct HashLayerProof {
  std::vector<Fr> eval_dim, eval_read, eval_final, eval_derefs;
  PolyEvalProof proof_ops, proof_mem;
  CombinedTableEvalProof proof_derefs;

  static HashLayerProof prove(const std::vector<Fr>& rand_mem, const std::vector<Fr>& rand_ops,
                              const DensifiedRepresentation& dense, const Subtables& subtables,
                              const SparsePolyCommitmentGens& gens, Transcript& transcript, RandomTape& tape) {
    SpanTimer st("HashLayer.prove");
    transcript.append_protocol_name("Lasso HashLayerProof");
    HashLayerProof out;
    for (auto& p : subtables.lookup_polys) out.eval_derefs.push_back(p.evaluate(rand_ops));
    out.proof_derefs = CombinedTableEvalProof::prove(subtables.combined_poly, out.eval_derefs, rand_ops,
                                                     gens.gens_derefs, transcript, tape);
    for (size_t i = 0; i < dense.C; i++) out.eval_dim.push_back(dense.dim[i].evaluate(rand_ops));
    for (size_t i = 0; i < dense.C; i++) out.eval_read.push_back(dense.read[i].evaluate(rand_ops));
    for (size_t i = 0; i < dense.C; i++) out.eval_final.push_back(dense.final_[i].evaluate(rand_mem));
    std::vector<Fr> evals_ops = out.eval_dim;
    evals_ops.insert(evals_ops.end(), out.eval_read.begin(), out.eval_read.end());
    evals_ops.resize(next_power_of_two(evals_ops.size()), Fr::zero());
    transcript.append_scalars("claim_evals_ops", evals_ops);
    std::vector<Fr> challenges_ops = transcript.challenge_vector("challenge_combine_n_to_one", log_2(evals_ops.size()));
    DensePolynomial poly_evals_ops(evals_ops);
    for (size_t i = challenges_ops.size(); i-- > 0;) poly_evals_ops.bound_poly_var_bot(challenges_ops[i]);
    Fr joint_claim_eval_ops = poly_evals_ops[0];
    std::vector<Fr> r_joint_ops = challenges_ops;
    r_joint_ops.insert(r_joint_ops.end(), rand_ops.begin(), rand_ops.end());
    transcript.append_scalar("joint_claim_eval_ops", joint_claim_eval_ops);
    out.proof_ops = PolyEvalProof::prove(dense.combined_l_variate_polys, r_joint_ops, joint_claim_eval_ops,
                                         gens.gens_combined_l_variate, transcript, tape);
    transcript.append_scalars("claim_evals_mem", out.eval_final);
    std::vector<Fr> challenges_mem =
        transcript.challenge_vector("challenge_combine_two_to_one", log_2(out.eval_final.size()));
    DensePolynomial poly_evals_mem = DensePolynomial::new_padded(out.eval_final);
    for (size_t i = challenges_mem.size(); i-- > 0;) poly_evals_mem.bound_poly_var_bot(challenges_mem[i]);
    Fr joint_claim_eval_mem = poly_evals_mem[0];
    std::vector<Fr> r_joint_mem = challenges_mem;
    r_joint_mem.insert(r_joint_mem.end(), rand_mem.begin(), rand_mem.end());
    transcript.append_scalar("joint_claim_eval_mem", joint_claim_eval_mem);
    out.proof_mem = PolyEvalProof::prove(dense.combined_log_m_variate_polys, r_joint_mem, joint_claim_eval_mem,
                                         gens.gens_combined_log_m_variate, transcript, tape);
    return out;
  }

  bool verify(const Strategy& S, const std::vector<Fr>& rand_mem, const std::vector<Fr>& rand_ops,
              const std::vector<std::array<Fr, 4>>& claims, const SparsePolynomialCommitment& comm,
              const SparsePolyCommitmentGens& gens, const PolyCommitment& comm_derefs, const Fr& gamma,
              const Fr& tau, Transcript& transcript) const {
    transcript.append_protocol_name("Lasso HashLayerProof");
    if (!proof_derefs.verify(rand_ops, eval_derefs, gens.gens_derefs, comm_derefs, transcript)) return false;
    std::vector<Fr> evals_ops = eval_dim;
    evals_ops.insert(evals_ops.end(), eval_read.begin(), eval_read.end());
    evals_ops.resize(next_power_of_two(evals_ops.size()), Fr::zero());
    transcript.append_scalars("claim_evals_ops", evals_ops);
    std::vector<Fr> challenges_ops = transcript.challenge_vector("challenge_combine_n_to_one", log_2(evals_ops.size()));
    DensePolynomial poly_evals_ops(evals_ops);
    for (size_t i = challenges_ops.size(); i-- > 0;) poly_evals_ops.bound_poly_var_bot(challenges_ops[i]);
    Fr joint_claim_eval_ops = poly_evals_ops[0];
    std::vector<Fr> r_joint_ops = challenges_ops;
    r_joint_ops.insert(r_joint_ops.end(), rand_ops.begin(), rand_ops.end());
    transcript.append_scalar("joint_claim_eval_ops", joint_claim_eval_ops);
    if (!proof_ops.verify_plain(gens.gens_combined_l_variate, transcript, r_joint_ops, joint_claim_eval_ops,
                                comm.l_variate_polys_commitment))
      return false;
    transcript.append_scalars("claim_evals_mem", eval_final);
    std::vector<Fr> challenges_mem = transcript.challenge_vector("challenge_combine_two_to_one", log_2(eval_final.size()));
    DensePolynomial poly_evals_mem = DensePolynomial::new_padded(eval_final);
    for (size_t i = challenges_mem.size(); i-- > 0;) poly_evals_mem.bound_poly_var_bot(challenges_mem[i]);
    Fr joint_claim_eval_mem = poly_evals_mem[0];
    std::vector<Fr> r_joint_mem = challenges_mem;
    r_joint_mem.insert(r_joint_mem.end(), rand_mem.begin(), rand_mem.end());
    transcript.append_scalar("joint_claim_eval_mem", joint_claim_eval_mem);
    if (!proof_mem.verify_plain(gens.gens_combined_log_m_variate, transcript, r_joint_mem, joint_claim_eval_mem,
                                comm.log_m_variate_polys_commitment))
      return false;
    // check_reed_solomon_fingerprints (memory_checking.rs:477-523)
    Fr init_addr = identity_poly_evaluate(rand_mem);
    Fr g2 = gamma.square();
    auto hash_func = [&](const Fr& a, const Fr& v, const Fr& t) { return t * g2 + v * gamma + a - tau; };
    for (size_t i = 0; i < claims.size(); i++) {
      size_t j = S.memory_to_dimension_index(i), k = S.memory_to_subtable_index(i);
      Fr init_memory = S.evaluate_subtable_mle(k, rand_mem);
      if (hash_func(init_addr, init_memory, Fr::zero()) != claims[i][0]) return false;
      if (hash_func(eval_dim[j], eval_derefs[i], eval_read[j]) != claims[i][1]) return false;
      if (hash_func(eval_dim[j], eval_derefs[i], eval_read[j] + Fr::one()) != claims[i][2]) return false;
      if (hash_func(init_addr, init_memory, eval_final[j]) != claims[i][3]) return false;
    }
    return true;
  }
};

// memory_checking.rs:26-147
struct MemoryCheckingProof {
  ProductLayerProof proof_prod_layer;
  HashLayerProof proof_hash_layer;

  static MemoryCheckingProof prove(const DensifiedRepresentation& dense, const Fr& gamma, const Fr& tau,
                                   const Subtables& subtables, const SparsePolyCommitmentGens& gens,
                                   Transcript& transcript, RandomTape& tape) {
    SpanTimer st("MemoryChecking.prove");
    transcript.append_protocol_name("Lasso MemoryCheckingProof");
    std::vector<GrandProducts> gps = subtables.to_grand_products(dense, gamma, tau);
    std::vector<Fr> rand_mem, rand_ops;
    MemoryCheckingProof out;
    out.proof_prod_layer = ProductLayerProof::prove(gps, transcript, rand_mem, rand_ops);
    out.proof_hash_layer = HashLayerProof::prove(rand_mem, rand_ops, dense, subtables, gens, transcript, tape);
    return out;
  }
  bool verify(const Strategy& S, const SparsePolynomialCommitment& comm, const PolyCommitment& comm_derefs,
              const SparsePolyCommitmentGens& gens, const Fr& gamma, const Fr& tau, size_t s,
              Transcript& transcript) const {
    transcript.append_protocol_name("Lasso MemoryCheckingProof");
    size_t num_ops = next_power_of_two(s), num_cells = comm.m;
    std::vector<Fr> claims_mem, rand_mem, claims_ops, rand_ops;
    if (!proof_prod_layer.verify(num_ops, num_cells, transcript, claims_mem, rand_mem, claims_ops, rand_ops))
      return false;
    std::vector<std::array<Fr, 4>> claims;
    for (size_t i = 0; i < S.num_memories(); i++)
      claims.push_back({claims_mem[2 * i], claims_ops[2 * i], claims_ops[2 * i + 1], claims_mem[2 * i + 1]});
    return proof_hash_layer.verify(S, rand_mem, rand_ops, claims, comm, gens, comm_derefs, gamma, tau, transcript);
  }
};

// ---------------------------------------------------------------- lasso/surge.rs:84-275
struct SparsePolynomialEvaluationProof {
  PolyCommitment comm_derefs;  // CombinedTableCommitment { comm_ops_val }
  // PrimarySumcheck
  SumcheckInstanceProof primary_proof;
  Fr claimed_evaluation;
  std::vector<Fr> eval_derefs;
  CombinedTableEvalProof proof_derefs;
  MemoryCheckingProof memory_check;

  // surge.rs:118-211
  static SparsePolynomialEvaluationProof prove(const Strategy& S, DensifiedRepresentation& dense,
                                               const std::vector<Fr>& r, const SparsePolyCommitmentGens& gens,
                                               Transcript& transcript, RandomTape& tape) {
    SpanTimer st("SparsePoly.prove");
    transcript.append_protocol_name("Lasso SparsePolynomialEvaluationProof");
    if (r.size() != ark_log2(dense.s)) throw std::runtime_error("r.len() != log2(s) (surge.rs:131)");
    SparsePolynomialEvaluationProof out;
    Subtables subtables(S, dense.dim_usize, dense.s);
    out.comm_derefs = subtables.commit(gens.gens_derefs);
    append_combined_table_commitment(out.comm_derefs, "comm_poly_row_col_ops_val", transcript);
    EqPolynomial eq(r);
    out.claimed_evaluation = subtables.compute_sumcheck_claim(eq);
    transcript.append_scalar("claim_eval_scalar_product", out.claimed_evaluation);
    std::vector<DensePolynomial> combined;
    for (auto& p : subtables.lookup_polys) combined.push_back(p.clone());
    combined.emplace_back(eq.evals());
    std::vector<Fr> r_z, final_evals;
    out.primary_proof = SumcheckInstanceProof::prove_arbitrary(
        log_2(dense.s), combined, [&](const Fr* v) { return S.combine_lookups_eq(v); }, S.sumcheck_poly_degree(),
        transcript, r_z, final_evals);
    for (auto& p : subtables.lookup_polys) out.eval_derefs.push_back(p.evaluate(r_z));
    out.proof_derefs = CombinedTableEvalProof::prove(subtables.combined_poly, out.eval_derefs, r_z, gens.gens_derefs,
                                                     transcript, tape);
    std::vector<Fr> r_hash = transcript.challenge_vector("challenge_r_hash", 2);
    out.memory_check = MemoryCheckingProof::prove(dense, r_hash[0], r_hash[1], subtables, gens, transcript, tape);
    return out;
  }

  // surge.rs:213-271
  bool verify(const Strategy& S, const SparsePolynomialCommitment& commitment, const std::vector<Fr>& eq_randomness,
              const SparsePolyCommitmentGens& gens, Transcript& transcript) const {
    transcript.append_protocol_name("Lasso SparsePolynomialEvaluationProof");
    append_combined_table_commitment(comm_derefs, "comm_poly_row_col_ops_val", transcript);
    transcript.append_scalar("claim_eval_scalar_product", claimed_evaluation);
    Fr claim_last;
    std::vector<Fr> r_z;
    if (!primary_proof.verify(claimed_evaluation, log_2(commitment.s), S.sumcheck_poly_degree(), transcript,
                              claim_last, r_z))
      return false;
    Fr eq_eval = EqPolynomial(eq_randomness).evaluate(r_z);
    if (eval_derefs.size() != S.num_memories()) return false;
    if (eq_eval * S.combine_lookups(eval_derefs.data()) != claim_last) return false;
    if (!proof_derefs.verify(r_z, eval_derefs, gens.gens_derefs, comm_derefs, transcript)) return false;
    std::vector<Fr> r_mem_check = transcript.challenge_vector("challenge_r_hash", 2);
    return memory_check.verify(S, commitment, comm_derefs, gens, r_mem_check[0], r_mem_check[1], commitment.s,
                               transcript);
  }
};

// ---------------------------------------------------------------- ark-serialize (compressed) of the proof
struct ByteWriter {
  std::vector<uint8_t> b;
  void u64(uint64_t v) {
    for (int i = 0; i < 8; i++) b.push_back((uint8_t)(v >> (8 * i)));
  }
  void fr(const Fr& f) {
    uint8_t t[32];
    f.to_bytes(t);
    b.insert(b.end(), t, t + 32);
  }
  void pt(const Point& p) {
    uint8_t t[32];
    p.compress(t);
    b.insert(b.end(), t, t + 32);
  }
  void vec_fr(const std::vector<Fr>& v) {
    u64(v.size());
    for (auto& f : v) fr(f);
  }
  void arr_fr(const std::vector<Fr>& v) {
    for (auto& f : v) fr(f);
  }
  void vec_pt(const std::vector<Point>& v) {
    u64(v.size());
    for (auto& p : v) pt(p);
  }
};
inline void ser(ByteWriter& w, const SumcheckInstanceProof& p) {
  w.u64(p.compressed_polys.size());
  for (auto& c : p.compressed_polys) w.vec_fr(c.coeffs_except_linear_term);
}
inline void ser(ByteWriter& w, const DotProductProofLog& p) {
  w.vec_pt(p.bullet_reduction_proof.L_vec);
  w.vec_pt(p.bullet_reduction_proof.R_vec);
  w.pt(p.delta);
  w.pt(p.beta);
  w.fr(p.z1);
  w.fr(p.z2);
}
inline void ser(ByteWriter& w, const BatchedGrandProductArgument& p) {
  w.u64(p.proof.size());
  for (auto& l : p.proof) {
    ser(w, l.proof);
    w.vec_fr(l.claims_prod_left);
    w.vec_fr(l.claims_prod_right);
  }
}
// field order: surge.rs:92-104, 84-90; memory_checking.rs:26-37, 655-660, 313-329
inline std::vector<uint8_t> serialize_proof(const SparsePolynomialEvaluationProof& p) {
  ByteWriter w;
  w.vec_pt(p.comm_derefs.C);
  ser(w, p.primary_proof);
  w.fr(p.claimed_evaluation);
  w.arr_fr(p.eval_derefs);
  ser(w, p.proof_derefs.proof_table_eval.proof);
  const auto& pl = p.memory_check.proof_prod_layer;
  for (auto& e : pl.grand_product_evals)
    for (int k = 0; k < 4; k++) w.fr(e[k]);
  ser(w, pl.proof_mem);
  ser(w, pl.proof_ops);
  const auto& hl = p.memory_check.proof_hash_layer;
  w.arr_fr(hl.eval_dim);
  w.arr_fr(hl.eval_read);
  w.arr_fr(hl.eval_final);
  w.arr_fr(hl.eval_derefs);
  ser(w, hl.proof_ops.proof);
  ser(w, hl.proof_mem.proof);
  ser(w, hl.proof_derefs.proof_table_eval.proof);
  return w.b;
}
inline std::vector<uint8_t> serialize_commitment(const SparsePolynomialCommitment& c) {  // surge.rs:61-68
  ByteWriter w;
  w.vec_pt(c.l_variate_polys_commitment.C);
  w.vec_pt(c.log_m_variate_polys_commitment.C);
  w.u64(c.s);
  w.u64(c.log_m);
  w.u64(c.m);
  return w.b;
}

}  // namespace oracle
