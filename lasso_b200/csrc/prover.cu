// lasso_b200 — the host prover: mirrors the reference's
//   DensifiedRepresentation::from_lookup_indices / commit      (src/lasso/densified.rs:21-96)
//   SparsePolynomialEvaluationProof::prove                      (src/lasso/surge.rs:118-211)
//   MemoryCheckingProof / ProductLayerProof / HashLayerProof    (src/lasso/memory_checking.rs)
//   BatchedGrandProductArgument::prove                          (src/subprotocols/grand_product.rs:100-201)
//   SumcheckInstanceProof::{prove_arbitrary, prove_cubic_batched} (src/subprotocols/sumcheck.rs)
//   PolyEvalProof / DotProductProofLog / BulletReductionProof   (src/poly/dense_mlpoly.rs:301-359,
//                                                                src/subprotocols/{dot_product,bullet}.rs)
// with every field/curve loop on the GPU and only the Fiat–Shamir transcript, the round-polynomial
// interpolation and O(log n)-sized vector glue on the host.  One host<->device round trip per sumcheck
// round ((deg+1) x 32 B down, the challenge travels as a kernel argument).
//
// Bulletproofs on a GPU (bullet.rs:73-142): the reference folds the generator vector every round,
// G_L[i] <- u^-1 G_L[i] + u G_R[i] — 2n serial variable-base scalar multiplications per opening.  Here the
// generators are never folded: round k's L and R are MSMs over the ORIGINAL generators with scalars
// a[i] * W_k[t] (W_k = the 2^k products of u_r^{+-1}), so every group operation of the proof is a row-MSM
// over one fixed table T[w][j] = 2^(8w) G_j.  The group elements are identical; only the schedule differs.
#include "prover.cuh"

#include <thread>

#include "host_fq64.hpp"

namespace lb {

unsigned long long g_launches = 0;

// ---------------------------------------------------------------------------------------------- context
Ctx* ctx_create(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    throw std::runtime_error("lasso_b200 needs a CUDA device (sm_100a); there is no CPU fallback");
  if (device < 0 || device >= count) throw std::runtime_error("invalid device id");
  LB_CUDA_CHECK(cudaSetDevice(device));
  std::unique_ptr<Ctx> c(new Ctx());
  c->device = device;
  LB_CUDA_CHECK(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
  cudaMemPool_t pool;
  LB_CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t thr = UINT64_MAX;
  LB_CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  c->h_pin_bytes = 8u << 20;
  LB_CUDA_CHECK(cudaMallocHost((void**)&c->h_pin, c->h_pin_bytes));
  c->partial_elems = (size_t)bound_max_chunks() * 16384 + 65536;
  LB_CUDA_CHECK(cudaMalloc((void**)&c->d_partial, c->partial_elems * sizeof(fr_t)));
  c->small_elems = 65536;
  LB_CUDA_CHECK(cudaMalloc((void**)&c->d_small, c->small_elems * sizeof(fr_t)));
  LB_CUDA_CHECK(cudaMalloc((void**)&c->d_eq_scratch, (size_t)(4096 + (1 << 17) + 4096) * sizeof(fr_t)));
  LB_CUDA_CHECK(cudaMalloc((void**)&c->d_flag, 64));
  const char* sp = getenv("LASSO_B200_SPANS");
  c->span_sync = sp && sp[0] == '1';
  return c.release();
}
void ctx_destroy(Ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->st);
  cudaFree(c->d_partial);
  cudaFree(c->d_small);
  cudaFree(c->d_eq_scratch);
  cudaFree(c->d_flag);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  cudaFreeHost(c->h_pin);
  cudaStreamDestroy(c->st);
  delete c;
}

static FrVec to_frvec(const std::vector<fr_t>& v, size_t off, size_t n) {
  if (n > 32) throw std::runtime_error("challenge vector too long");
  FrVec f;
  for (size_t i = 0; i < n; i++) f.v[i] = v[off + i];
  return f;
}
// eq(r) table on the device (eq_poly.rs:21-38)
static void eq_evals_dev(Ctx* c, const std::vector<fr_t>& r, size_t off, size_t ell, fr_t* out) {
  launch_eq_evals(to_frvec(r, off, ell), (int)ell, out, c->d_eq_scratch, c->st);
  g_launches += ell <= 11 ? 1 : 3;
}

// ---------------------------------------------------------------------------------------------- generators
size_t gens_points_needed(size_t c, size_t s, size_t num_memories, size_t log_m) {
  size_t nv_l = log2_exact_or_ceil(next_pow2(2 * c * s));
  size_t nv_m = log2_exact_or_ceil(next_pow2(c)) + log_m;
  size_t nv_d = log2_exact_or_ceil(next_pow2(num_memories * s));
  size_t mx = std::max(nv_l, std::max(nv_m, nv_d));
  return ((size_t)1 << (mx - mx / 2)) + 2;
}
Gens* gens_create(Ctx* c, const uint64_t* stream_affine, size_t n_points, size_t cc, size_t s, size_t num_memories,
                  size_t log_m) {
  if (n_points < gens_points_needed(cc, s, num_memories, log_m)) return nullptr;
  std::unique_ptr<Gens> g(new Gens());
  g->ctx = c;
  g->n_points = n_points;
  g->c = cc;
  g->s = s;
  g->num_memories = num_memories;
  g->log_m = log_m;
  g->nv_l = log2_exact_or_ceil(next_pow2(2 * cc * s));
  g->nv_m = log2_exact_or_ceil(next_pow2(cc)) + log_m;
  g->nv_d = log2_exact_or_ceil(next_pow2(num_memories * s));
  g->d_bases_ark.alloc(c, n_points * 2);
  LB_CUDA_CHECK(cudaMemcpyAsync(g->d_bases_ark.p, stream_affine, n_points * 64, cudaMemcpyHostToDevice, c->st));
  g->d_table.alloc(c, (size_t)kMsmFullWindows * n_points);
  launch_build_table(g->d_bases_ark.p, n_points, g->d_table.p, n_points, kMsmFullWindows, c->st);
  g_launches += kMsmFullWindows;
  c->sync();
  return g.release();
}

// ---------------------------------------------------------------------------------------------- MSM helpers
struct MsmOut {
  std::vector<uint8_t> comp;  // 32 B per row
};
// rows of u32 integer scalars over generator columns [0, ncols)
static std::vector<uint8_t> msm_rows_u32(Ctx* c, const Gens& g, const uint32_t* d_scal, size_t row_stride, int nrows,
                                         int ncols, unsigned max_bits) {
  int nw = msm_windows_for_bits(max_bits);
  if (nw > 5) throw std::runtime_error("u32 MSM path: scalars wider than 32 bits");
  DBuf<pt_ext> part(c, msm_partials_count(nrows, ncols, nw));
  DBuf<uint32_t> comp(c, (size_t)nrows * 8);
  launch_msm_rows(g.d_table.p, g.n_points, 1, d_scal, 1, row_stride, nrows, ncols, nw, part.p, nullptr, comp.p, nullptr,
                  c->st);
  g_launches += 2;
  std::vector<uint8_t> out((size_t)nrows * 32);
  c->d2h(out.data(), comp.p, out.size());
  return out;
}
// rows of Montgomery Fr scalars (device) over generator columns [col0, col0 + ncols); full-width windows
static std::vector<uint8_t> msm_rows_fr(Ctx* c, const Gens& g, const fr_t* d_scal_mont, int nrows, int ncols,
                                        size_t col0) {
  DBuf<fr_t> canon(c, (size_t)nrows * ncols);
  launch_canonicalize(d_scal_mont, canon.p, (size_t)nrows * ncols, c->d_flag, c->st);
  int nw = kMsmFullWindows;
  DBuf<pt_ext> part(c, msm_partials_count(nrows, ncols, nw));
  std::vector<uint8_t> out((size_t)nrows * 32);
  if (nrows <= 8) {
    // a couple of points per Bulletproofs round: ship (X, Y, Z) and invert on the host (3 us vs ~100 us
    // for the same serial chain on one GPU thread)
    DBuf<uint32_t> raw(c, (size_t)nrows * 24);
    launch_msm_rows(g.d_table.p + col0, g.n_points, 1, canon.p, 8, (size_t)ncols, nrows, ncols, nw, part.p, nullptr,
                    nullptr, raw.p, c->st);
    g_launches += 3;
    uint32_t xyz[8 * 24];
    c->d2h(xyz, raw.p, (size_t)nrows * 96);
    for (int i = 0; i < nrows; i++) h64::compress_xyz(xyz + 24 * i, out.data() + 32 * i);
    return out;
  }
  DBuf<uint32_t> comp(c, (size_t)nrows * 8);
  launch_msm_rows(g.d_table.p + col0, g.n_points, 1, canon.p, 8, (size_t)ncols, nrows, ncols, nw, part.p, nullptr, comp.p,
                  nullptr, c->st);
  g_launches += 3;
  c->d2h(out.data(), comp.p, out.size());
  return out;
}

// DensePolynomial::commit (dense_mlpoly.rs:152-181) for an integer-valued polynomial of 2^nv entries
static std::vector<uint8_t> commit_u32(Ctx* c, const Gens& g, const uint32_t* d_vals, size_t nv, unsigned max_bits) {
  size_t L = (size_t)1 << (nv / 2), R = (size_t)1 << (nv - nv / 2);
  if (R + 2 > g.n_points) throw std::runtime_error("generator stream too short for this polynomial");
  return msm_rows_u32(c, g, d_vals, R, (int)L, (int)R, max_bits);
}

// ---------------------------------------------------------------------------------------------- densify
Dense* densify(Ctx* c, const uint64_t* indices, size_t n, size_t C, size_t log_m, int* err) {
  SpanTimer sp(c, "Densify");
  *err = 0;
  if (n == 0 || C == 0 || C > 16 || log_m < 1 || log_m > 28) {
    *err = 4;
    return nullptr;
  }
  std::unique_ptr<Dense> d(new Dense());
  d->ctx = c;
  d->C = C;
  d->s = next_pow2(n);
  d->log_m = log_m;
  d->m = (size_t)1 << log_m;
  d->nv_l = log2_exact_or_ceil(next_pow2(2 * C * d->s));
  d->nv_m = log2_exact_or_ceil(next_pow2(C)) + log_m;
  const size_t s = d->s, m = d->m, nl = (size_t)1 << d->nv_l, nm = (size_t)1 << d->nv_m;
  // pinned, reused across calls: no per-call page faults, and the upload runs at full PCIe rate
  uint32_t* l_host = c->stage(nl + nm);
  uint32_t* m_host = l_host + nl;
  if (nl > 2 * C * s) memset(l_host + 2 * C * s, 0, (nl - 2 * C * s) * sizeof(uint32_t));
  memset(m_host, 0, nm * sizeof(uint32_t));
  // densified.rs:33-56: per dimension, pad with address 0 and run the (inherently sequential) timestamp
  // counters; dimensions are independent, so one host thread each.
  std::vector<int> bad(C, 0);
  auto work = [&](size_t i) {
    uint32_t* dim = l_host + i * s;
    uint32_t* rd = l_host + (C + i) * s;
    uint32_t* fin = m_host + i * m;
    for (size_t k = 0; k < s; k++) {
      uint64_t addr = k < n ? indices[k * C + i] : 0;
      if (addr >= m) {
        bad[i] = 1;
        return;
      }
      dim[k] = (uint32_t)addr;
      uint32_t ts = fin[addr];
      rd[k] = ts;
      fin[addr] = ts + 1;
    }
  };
  {
    std::vector<std::thread> th;
    for (size_t i = 1; i < C; i++) th.emplace_back(work, i);
    work(0);
    for (auto& t : th) t.join();
  }
  for (size_t i = 0; i < C; i++)
    if (bad[i]) {
      *err = 3;
      return nullptr;
    }
  d->d_l_u32.alloc(c, nl);
  d->d_m_u32.alloc(c, nm);
  d->d_l_fr.alloc(c, nl);
  d->d_m_fr.alloc(c, nm);
  LB_CUDA_CHECK(cudaMemcpyAsync(d->d_l_u32.p, l_host, nl * 4, cudaMemcpyHostToDevice, c->st));
  LB_CUDA_CHECK(cudaMemcpyAsync(d->d_m_u32.p, m_host, nm * 4, cudaMemcpyHostToDevice, c->st));
  launch_from_u32(d->d_l_u32.p, d->d_l_fr.p, nl, c->st);  // DensePolynomial::from_usize + merge
  launch_from_u32(d->d_m_u32.p, d->d_m_fr.p, nm, c->st);
  g_launches += 2;
  c->sync();
  return d.release();
}

// densified.rs:77-96 -> serialised SparsePolynomialCommitment (surge.rs:61-68)
std::vector<uint8_t> commit(Ctx* c, const Dense& d, const Gens& g) {
  SpanTimer sp(c, "DensifiedRepresentation.commit");
  if (g.nv_l != d.nv_l || g.nv_m != d.nv_m) throw std::runtime_error("generators were built for different (c, s, log_m)");
  unsigned bits = (unsigned)std::max(d.log_m, (size_t)(log2_exact_or_ceil(d.s) + 1));
  ByteWriter w;
  w.vec_pts(commit_u32(c, g, d.d_l_u32.p, d.nv_l, bits));
  w.vec_pts(commit_u32(c, g, d.d_m_u32.p, d.nv_m, bits));
  w.u64(d.s);
  w.u64(d.log_m);
  w.u64(d.m);
  return w.b;
}

// ---------------------------------------------------------------------------------------------- UniPoly
// unipoly.rs:30-54: coefficients of the polynomial through (0, e_0) .. (n-1, e_{n-1}).  The solution of the
// Vandermonde system is unique, so it is computed with a cached inverse matrix instead of eliminating
// per round.
static const std::vector<fr_t>& inv_vandermonde(size_t n) {
  static std::map<size_t, std::vector<fr_t>> cache;
  auto it = cache.find(n);
  if (it != cache.end()) return it->second;
  std::vector<fr_t> a(n * 2 * n, fr_zero());  // [V | I], Gauss-Jordan
  for (size_t i = 0; i < n; i++) {
    fr_t x = fr_from_u64(i), p = fr_one();
    for (size_t j = 0; j < n; j++) {
      a[i * 2 * n + j] = p;
      p = fr_mul(p, x);
    }
    a[i * 2 * n + n + i] = fr_one();
  }
  for (size_t col = 0; col < n; col++) {
    size_t piv = col;
    while (fr_is_zero(a[piv * 2 * n + col])) piv++;
    if (piv != col)
      for (size_t k = 0; k < 2 * n; k++) std::swap(a[piv * 2 * n + k], a[col * 2 * n + k]);
    fr_t inv = fr_inv(a[col * 2 * n + col]);
    for (size_t k = 0; k < 2 * n; k++) a[col * 2 * n + k] = fr_mul(a[col * 2 * n + k], inv);
    for (size_t row = 0; row < n; row++) {
      if (row == col) continue;
      fr_t f = a[row * 2 * n + col];
      if (fr_is_zero(f)) continue;
      for (size_t k = 0; k < 2 * n; k++) a[row * 2 * n + k] = fr_sub(a[row * 2 * n + k], fr_mul(f, a[col * 2 * n + k]));
    }
  }
  std::vector<fr_t> inv(n * n);
  for (size_t i = 0; i < n; i++)
    for (size_t j = 0; j < n; j++) inv[i * n + j] = a[i * 2 * n + n + j];
  return cache[n] = inv;
}
static std::vector<fr_t> unipoly_from_evals(const std::vector<fr_t>& evals) {
  size_t n = evals.size();
  const std::vector<fr_t>& inv = inv_vandermonde(n);
  std::vector<fr_t> coeffs(n, fr_zero());
  for (size_t i = 0; i < n; i++)
    for (size_t j = 0; j < n; j++) coeffs[i] = fr_add(coeffs[i], fr_mul(inv[i * n + j], evals[j]));
  return coeffs;
}
static fr_t unipoly_evaluate(const std::vector<fr_t>& coeffs, const fr_t& r) {  // unipoly.rs:72-80
  fr_t eval = coeffs[0], power = r;
  for (size_t i = 1; i < coeffs.size(); i++) {
    eval = fr_add(eval, fr_mul(power, coeffs[i]));
    power = fr_mul(power, r);
  }
  return eval;
}
static void unipoly_append(const std::vector<fr_t>& coeffs, Transcript& t) {  // unipoly.rs:112-120
  t.append_message("poly", std::string("UniPoly_begin"));
  for (auto& cf : coeffs) t.append_scalar("coeff", cf);
  t.append_message("poly", std::string("UniPoly_end"));
}
typedef std::vector<std::vector<fr_t>> SumcheckProof;  // compressed polys: coeffs without the linear term
static std::vector<fr_t> unipoly_compress(const std::vector<fr_t>& coeffs) {  // unipoly.rs:82-88
  std::vector<fr_t> c;
  c.push_back(coeffs[0]);
  for (size_t i = 2; i < coeffs.size(); i++) c.push_back(coeffs[i]);
  return c;
}
static void ser_sumcheck(ByteWriter& w, const SumcheckProof& p) {
  w.u64(p.size());
  for (auto& c : p) w.vec_fr(c);
}

// ---------------------------------------------------------------------------------------------- sumcheck
// sumcheck.rs:149-260 over device polynomials W_k = base + k*stride (k <= alpha, the last is eq)
static SumcheckProof prove_arbitrary(Ctx* c, const Strategy& S, fr_t* base, size_t stride, size_t len,
                                     Transcript& transcript, std::vector<fr_t>& r) {
  SpanTimer sp(c, "Sumcheck.prove");
  SumcheckProof proof;
  r.clear();
  const int npts = S.sumcheck_poly_degree() + 1, npolys = S.num_memories() + 1;
  std::vector<fr_t> evals(npts);
  while (len > 1) {
    size_t half = len / 2;
    launch_sumcheck_eval_arbitrary(S, base, stride, half, c->d_partial, c->d_small, c->st);
    g_launches += 2;
    c->d2h(evals.data(), c->d_small, (size_t)npts * sizeof(fr_t));
    std::vector<fr_t> coeffs = unipoly_from_evals(evals);
    unipoly_append(coeffs, transcript);
    fr_t r_j = transcript.challenge_scalar("challenge_nextround");
    r.push_back(r_j);
    launch_bind_top(base, stride, npolys, half, r_j, c->st);
    g_launches += 1;
    proof.push_back(unipoly_compress(coeffs));
    len = half;
  }
  return proof;
}

// ---------------------------------------------------------------------------------------------- grand products
// GrandProductCircuit (grand_product.rs:14-66): layer k is one contiguous array of N/2^k elements,
// left_vec[k] = first half, right_vec[k] = second half; layer k+1[i] = layer k[i] * layer k[i + N/2^(k+1)].
struct Circuit {
  DBuf<fr_t> tree;  // 2N elements: layer 0 at 0, layer 1 at N, layer 2 at N + N/2, ...
  size_t N = 0, num_layers = 0;
  fr_t* layer(size_t k) const {
    size_t off = 0, len = N;
    for (size_t i = 0; i < k; i++) {
      off += len;
      len /= 2;
    }
    return tree.p + off;
  }
  size_t layer_len(size_t k) const { return N >> k; }
};
static void build_tree(Ctx* c, Circuit& ci) {  // grand_product.rs:38-58 (layer 0 already filled)
  ci.num_layers = log2_exact_or_ceil(ci.N);
  for (size_t k = 0; k + 1 < ci.num_layers; k++) {
    launch_product_layer(ci.layer(k), ci.layer(k + 1), ci.layer_len(k + 1), c->st);
    g_launches += 1;
  }
}

struct LayerProof {
  SumcheckProof proof;
  std::vector<fr_t> claims_prod_left, claims_prod_right;
};
typedef std::vector<LayerProof> GPAProof;

// BatchedGrandProductArgument::prove (grand_product.rs:100-201) with prove_cubic_batched (sumcheck.rs:26-135)
static GPAProof prove_gpa(Ctx* c, std::vector<Circuit*>& circuits, std::vector<fr_t> claims_to_verify,
                          Transcript& transcript, std::vector<fr_t>& rand_out) {
  SpanTimer sp(c, "BatchedGrandProductArgument.prove");
  GPAProof out;
  const int ncirc = (int)circuits.size();
  const size_t num_layers = circuits[0]->num_layers;
  DBuf<fr_t*> d_A(c, ncirc), d_B(c, ncirc), d_AB(c, 2 * ncirc);
  DBuf<fr_t> eqbuf(c, std::max<size_t>(circuits[0]->N / 2, 1)), eqbuf2(c, std::max<size_t>(circuits[0]->N / 4, 1));
  std::vector<fr_t*> hA(ncirc), hB(ncirc), hAB(2 * ncirc);
  std::vector<fr_t> rand;
  std::vector<fr_t> ev((size_t)ncirc * 3), fin((size_t)2 * ncirc);
  for (size_t layer_id = num_layers; layer_id-- > 0;) {
    const size_t len = circuits[0]->layer_len(layer_id);
    size_t half_len = len / 2;  // |A| = |B| = |C|
    for (int k = 0; k < ncirc; k++) {
      hA[k] = circuits[k]->layer(layer_id);
      hB[k] = hA[k] + half_len;
      hAB[2 * k] = hA[k];
      hAB[2 * k + 1] = hB[k];
    }
    LB_CUDA_CHECK(cudaMemcpyAsync(d_A.p, hA.data(), ncirc * sizeof(fr_t*), cudaMemcpyHostToDevice, c->st));
    LB_CUDA_CHECK(cudaMemcpyAsync(d_B.p, hB.data(), ncirc * sizeof(fr_t*), cudaMemcpyHostToDevice, c->st));
    LB_CUDA_CHECK(cudaMemcpyAsync(d_AB.p, hAB.data(), 2 * ncirc * sizeof(fr_t*), cudaMemcpyHostToDevice, c->st));
    c->sync();  // the host arrays are reused next layer
    eq_evals_dev(c, rand, 0, rand.size(), eqbuf.p);  // poly_C = eq(rand), grand_product.rs:122
    std::vector<fr_t> coeff_vec = transcript.challenge_vector("rand_coeffs_next_layer", ncirc);
    fr_t e = fr_zero();
    for (int k = 0; k < ncirc; k++) e = fr_add(e, fr_mul(claims_to_verify[k], coeff_vec[k]));
    LayerProof lp;
    std::vector<fr_t> rand_prod;
    size_t cur = half_len;  // current length of A_k / B_k / C
    fr_t* Ccur = eqbuf.p;
    fr_t* Cnext = eqbuf2.p;
    if (cur > 1) {  // round 0 evaluation; later rounds come out of the fused bind+eval kernel
      launch_sumcheck_eval_cubic(d_A.p, d_B.p, Ccur, ncirc, cur / 2, c->d_partial, c->d_small, c->st);
      g_launches += 2;
    }
    while (cur > 1) {
      size_t half = cur / 2;
      c->d2h(ev.data(), c->d_small, ev.size() * sizeof(fr_t));
      fr_t c0 = fr_zero(), c2 = fr_zero(), c3 = fr_zero();
      for (int k = 0; k < ncirc; k++) {  // sumcheck.rs:95-97
        c0 = fr_add(c0, fr_mul(ev[3 * k], coeff_vec[k]));
        c2 = fr_add(c2, fr_mul(ev[3 * k + 1], coeff_vec[k]));
        c3 = fr_add(c3, fr_mul(ev[3 * k + 2], coeff_vec[k]));
      }
      std::vector<fr_t> evals = {c0, fr_sub(e, c0), c2, c3};  // eval(1) = e - eval(0), sumcheck.rs:99-104
      std::vector<fr_t> coeffs = unipoly_from_evals(evals);
      unipoly_append(coeffs, transcript);
      fr_t r_j = transcript.challenge_scalar("challenge_nextround");
      rand_prod.push_back(r_j);
      if (half > 1) {
        // bind with r_j and evaluate the next round in one pass (sumcheck.rs:116-120 + 63-89)
        g_launches += launch_sumcheck_bind_eval_cubic(d_A.p, d_B.p, Ccur, Cnext, ncirc, half, r_j, c->d_partial,
                                                      c->d_small, c->st);
        std::swap(Ccur, Cnext);
      } else {
        launch_bind_top_ptrs(d_AB.p, 2 * ncirc, half, r_j, c->st);
        g_launches += 1;
      }
      e = unipoly_evaluate(coeffs, r_j);
      lp.proof.push_back(unipoly_compress(coeffs));
      cur = half;
    }
    // claims_prod = (A_k[0], B_k[0]); gather the 2*ncirc scalars
    for (int k = 0; k < ncirc; k++) {
      LB_CUDA_CHECK(cudaMemcpyAsync(c->h_pin + (size_t)(2 * k) * 32, hA[k], 32, cudaMemcpyDeviceToHost, c->st));
      LB_CUDA_CHECK(cudaMemcpyAsync(c->h_pin + (size_t)(2 * k + 1) * 32, hB[k], 32, cudaMemcpyDeviceToHost, c->st));
    }
    c->sync();
    memcpy(fin.data(), c->h_pin, fin.size() * 32);
    for (int k = 0; k < ncirc; k++) {
      lp.claims_prod_left.push_back(fin[2 * k]);
      lp.claims_prod_right.push_back(fin[2 * k + 1]);
    }
    for (int k = 0; k < ncirc; k++) {
      transcript.append_scalar("claim_prod_left", lp.claims_prod_left[k]);
      transcript.append_scalar("claim_prod_right", lp.claims_prod_right[k]);
    }
    fr_t r_layer = transcript.challenge_scalar("challenge_r_layer");
    for (int k = 0; k < ncirc; k++)
      claims_to_verify[k] = fr_add(lp.claims_prod_left[k],
                                   fr_mul(r_layer, fr_sub(lp.claims_prod_right[k], lp.claims_prod_left[k])));
    std::vector<fr_t> ext = {r_layer};
    ext.insert(ext.end(), rand_prod.begin(), rand_prod.end());
    rand = ext;
    out.push_back(std::move(lp));
  }
  rand_out = rand;
  return out;
}
static void ser_gpa(ByteWriter& w, const GPAProof& p) {
  w.u64(p.size());
  for (auto& l : p) {
    ser_sumcheck(w, l.proof);
    w.vec_fr(l.claims_prod_left);
    w.vec_fr(l.claims_prod_right);
  }
}

// ---------------------------------------------------------------------------------------------- openings
struct DotProductProofLogBytes {  // dot_product.rs:152-159 field order
  std::vector<uint8_t> L_vec, R_vec;  // 32 B per point
  uint8_t delta[32], beta[32];
  fr_t z1, z2;
};
static void ser_dpl(ByteWriter& w, const DotProductProofLogBytes& p) {
  w.vec_pts(p.L_vec);
  w.vec_pts(p.R_vec);
  w.raw(p.delta, 32);
  w.raw(p.beta, 32);
  w.fr(p.z1);
  w.fr(p.z2);
}

__global__ void set_tail_kernel(fr_t* sL, fr_t* sR, size_t n, const fr_t* ip, fr_t blind_L, fr_t blind_R) {
  if (threadIdx.x || blockIdx.x) return;
  sL[n] = ip[0];      // c_L on Q
  sL[n + 1] = blind_L;  // on H
  sR[n] = ip[1];
  sR[n + 1] = blind_R;
}
__global__ void set_elems_kernel(fr_t* dst, fr_t a, fr_t b) {
  if (threadIdx.x || blockIdx.x) return;
  dst[0] = a;
  dst[1] = b;
}

// PolyEvalProof::prove (dense_mlpoly.rs:301-359) -> DotProductProofLog::prove (dot_product.rs:166-249)
// -> BulletReductionProof::prove (bullet.rs:40-154).  Z: device polynomial of 2^nv elements.
static DotProductProofLogBytes prove_poly_eval(Ctx* c, const Gens& g, const fr_t* Z, size_t nv,
                                               const std::vector<fr_t>& r, const fr_t& Zr, Transcript& transcript,
                                               RandomTape& tape) {
  SpanTimer sp(c, "DensePolyEval.prove");
  transcript.append_protocol_name("polynomial evaluation proof");
  if (r.size() != nv) throw std::runtime_error("PolyEvalProof: r.len() != num_vars");
  const size_t lv = nv / 2, rv = nv - nv / 2, L_size = (size_t)1 << lv, n = (size_t)1 << rv;  // n = R_size
  if (n + 2 > g.n_points) throw std::runtime_error("generator stream too short");
  const size_t lg_n = rv;
  // L, R = factored eq evals (eq_poly.rs:44-52); LZ = L . Z (dense_mlpoly.rs:183-207)
  std::unique_ptr<SpanTimer> sp1(new SpanTimer(c, "PE.1 eq+bound"));
  DBuf<fr_t> Lvec(c, L_size), a(c, n), b(c, n);
  eq_evals_dev(c, r, 0, lv, Lvec.p);
  eq_evals_dev(c, r, lv, rv, b.p);  // a_vec of the dot product proof = R
  if ((size_t)bound_max_chunks() * n > c->partial_elems) throw std::runtime_error("bound scratch too small");
  launch_bound(Z, Lvec.p, L_size, n, c->d_partial, a.p, c->st);  // x_vec = LZ
  g_launches += 2;

  // ---- DotProductProofLog::prove
  sp1.reset(new SpanTimer(c, "PE.2 Cx,Cy,append a"));
  transcript.append_protocol_name("dot product proof (log)");
  fr_t d = tape.random_scalar("d");
  fr_t r_delta = tape.random_scalar("r_delta");
  fr_t r_beta = tape.random_scalar("r_delta");  // sic (dot_product.rs:189)
  std::vector<fr_t> v1 = tape.random_vector("blinds_vec_1", 2 * lg_n);
  std::vector<fr_t> v2 = tape.random_vector("blinds_vec_2", 2 * lg_n);
  DotProductProofLogBytes out;
  // Cx = batch_commit(x_vec, blind_x = 0) ; Cy = y*Q + 0*h
  std::vector<uint8_t> Cx = msm_rows_fr(c, g, a.p, 1, (int)n, 0);
  transcript.append_point_compressed("Cx", Cx.data());
  DBuf<fr_t> two(c, 2);
  set_elems_kernel<<<1, 32, 0, c->st>>>(two.p, Zr, fr_zero());
  g_launches += 1;
  std::vector<uint8_t> Cy = msm_rows_fr(c, g, two.p, 1, 2, n);
  transcript.append_point_compressed("Cy", Cy.data());
  {  // append_scalars(b"a", a_vec): canonical bytes straight from the device
    DBuf<fr_t> canon(c, n);
    launch_canonicalize(b.p, canon.p, n, c->d_flag, c->st);
    g_launches += 1;
    std::vector<uint8_t> bytes(n * 32);
    c->d2h(bytes.data(), canon.p, bytes.size());
    transcript.append_scalars_bytes("a", bytes.data(), n);
  }
  // ---- BulletReductionProof::prove with unfolded generators (see file header)
  sp1.reset(new SpanTimer(c, "PE.3 bullet rounds"));
  fr_t blind_fin = fr_zero();  // blind_Gamma = blind_x + blind_y = 0
  DBuf<fr_t> W0(c, n), W1(c, n), sLR(c, 2 * (n + 2));
  fr_t* W = W0.p;
  fr_t* Wn = W1.p;
  set_elems_kernel<<<1, 32, 0, c->st>>>(W, fr_one(), fr_zero());
  g_launches += 1;
  fr_t* sL = sLR.p;
  fr_t* sR = sLR.p + (n + 2);
  size_t m = n, nw_count = 1;  // current vector length, number of weights
  for (size_t round = 0; m != 1; round++) {
    size_t h = m / 2;
    launch_cross_inner_products(a.p, b.p, h, c->d_partial, c->d_small, c->st);  // c_L, c_R (bullet.rs:78-79)
    launch_bullet_scalars(a.p, W, n, m, sL, sR, c->st);
    set_tail_kernel<<<1, 32, 0, c->st>>>(sL, sR, n, c->d_small, v1[round], v2[round]);
    g_launches += 4;
    std::vector<uint8_t> LR = msm_rows_fr(c, g, sLR.p, 2, (int)(n + 2), 0);
    transcript.append_point_compressed("L", LR.data());
    transcript.append_point_compressed("R", LR.data() + 32);
    fr_t u = transcript.challenge_scalar("u");
    fr_t u_inv = fr_inv(u);
    launch_fold_ab(a.p, b.p, h, u, u_inv, c->st);  // bullet.rs:127-130 (scalars only; G stays unfolded)
    launch_expand_weights(W, Wn, nw_count, u, u_inv, c->st);
    g_launches += 2;
    std::swap(W, Wn);
    nw_count *= 2;
    blind_fin = fr_add(blind_fin, fr_add(fr_mul(fr_mul(v1[round], u), u), fr_mul(fr_mul(v2[round], u_inv), u_inv)));
    out.L_vec.insert(out.L_vec.end(), LR.begin(), LR.begin() + 32);
    out.R_vec.insert(out.R_vec.end(), LR.begin() + 32, LR.begin() + 64);
    m = h;
  }
  sp1.reset(new SpanTimer(c, "PE.4 delta,beta"));
  fr_t ab[2];
  LB_CUDA_CHECK(cudaMemcpyAsync(c->h_pin, a.p, 32, cudaMemcpyDeviceToHost, c->st));
  LB_CUDA_CHECK(cudaMemcpyAsync(c->h_pin + 32, b.p, 32, cudaMemcpyDeviceToHost, c->st));
  c->sync();
  memcpy(ab, c->h_pin, 64);
  fr_t x_hat = ab[0], a_hat = ab[1], rhat_Gamma = blind_fin;
  fr_t y_hat = fr_mul(x_hat, a_hat);
  // delta = d * g_hat + r_delta * h with g_hat = sum_j W[j] G_j  (dot_product.rs:219-227)
  launch_scale(W, sL, n, d, c->st);
  set_elems_kernel<<<1, 32, 0, c->st>>>(sL + n, fr_zero(), r_delta);
  g_launches += 2;
  std::vector<uint8_t> delta = msm_rows_fr(c, g, sL, 1, (int)(n + 2), 0);
  memcpy(out.delta, delta.data(), 32);
  transcript.append_point_compressed("delta", out.delta);
  // beta = d * Q + r_beta * h  (dot_product.rs:229-230)
  set_elems_kernel<<<1, 32, 0, c->st>>>(two.p, d, r_beta);
  g_launches += 1;
  std::vector<uint8_t> beta = msm_rows_fr(c, g, two.p, 1, 2, n);
  memcpy(out.beta, beta.data(), 32);
  transcript.append_point_compressed("beta", out.beta);
  fr_t cc = transcript.challenge_scalar("c");
  out.z1 = fr_add(d, fr_mul(cc, y_hat));
  out.z2 = fr_add(fr_mul(a_hat, fr_add(fr_mul(cc, rhat_Gamma), r_beta)), r_delta);
  return out;
}

// CombinedTableEvalProof::prove (subtables/mod.rs:284-313 + prove_single 230-281) and the two analogous
// n-to-1 reductions of HashLayerProof::prove: fold `evals` with bound_poly_var_bot in reverse challenge order.
static DotProductProofLogBytes prove_joint(Ctx* c, const Gens& g, const fr_t* Z, size_t nv, std::vector<fr_t> evals,
                                           bool pad_before_append, const char* evals_label, const char* chal_label,
                                           const char* joint_label, const std::vector<fr_t>& r,
                                           Transcript& transcript, RandomTape& tape) {
  std::vector<fr_t> padded = evals;
  padded.resize(next_pow2(padded.size()), fr_zero());
  if (pad_before_append) evals = padded;
  transcript.append_scalars(evals_label, evals.data(), evals.size());
  std::vector<fr_t> challenges = transcript.challenge_vector(chal_label, log2_exact_or_ceil(evals.size()));
  std::vector<fr_t> pe = padded;
  for (size_t i = challenges.size(); i-- > 0;) {  // bound_poly_var_bot (dense_mlpoly.rs:218-225), tiny: host
    size_t half = pe.size() / 2;
    for (size_t k = 0; k < half; k++)
      pe[k] = fr_add(pe[2 * k], fr_mul(challenges[i], fr_sub(pe[2 * k + 1], pe[2 * k])));
    pe.resize(half);
  }
  fr_t joint = pe[0];
  std::vector<fr_t> r_joint = challenges;
  r_joint.insert(r_joint.end(), r.begin(), r.end());
  transcript.append_scalar(joint_label, joint);
  return prove_poly_eval(c, g, Z, nv, r_joint, joint, transcript, tape);
}

// ---------------------------------------------------------------------------------------------- prove
std::vector<uint8_t> prove(Ctx* c, const Strategy& S, Dense& dense, const std::vector<fr_t>& r, const Gens& g,
                           const std::string& transcript_label, const std::string& tape_label, const fr_t& tape_seed,
                           std::vector<fr_t>* challenges) {
  SpanTimer sp_all(c, "SparsePoly.prove");
  Transcript transcript(transcript_label);
  transcript.trace = challenges;
  RandomTape tape(tape_label, tape_seed);
  const size_t s = dense.s, C = dense.C, M = dense.m, alpha = (size_t)S.num_memories();
  const size_t log_s = log2_exact_or_ceil(s);
  if ((size_t)S.C != C || (size_t)S.log_m != dense.log_m) throw std::runtime_error("strategy does not match the densified representation");
  if (g.nv_d != log2_exact_or_ceil(next_pow2(alpha * s)) || g.nv_l != dense.nv_l || g.nv_m != dense.nv_m)
    throw std::runtime_error("generators were built for different (c, s, num_memories, log_m)");
  transcript.append_protocol_name("Lasso SparsePolynomialEvaluationProof");

  // ---- Subtables::new (subtables/mod.rs:116-129): materialise, gather, merge
  const size_t nv_d = g.nv_d, nd = (size_t)1 << nv_d;
  const int nsub = S.num_subtables();
  DBuf<fr_t> tables_fr(c, (size_t)nsub * M);
  DBuf<uint32_t> tables_u32(c, (size_t)nsub * M);
  DBuf<fr_t> E(c, nd);          // combined_poly = E_0 | .. | E_{alpha-1} | 0-pad
  DBuf<uint32_t> E_u32(c, nd);  // same values as integers for the small-scalar commit
  {
    SpanTimer sp(c, "Subtables.new");
    launch_materialize_subtables(S, tables_fr.p, tables_u32.p, c->st);
    launch_gather_lookup_polys(S, tables_fr.p, tables_u32.p, dense.nz(), s, E.p, s, E_u32.p, c->st);
    g_launches += 2;
    if (nd > alpha * s) {
      launch_fill_zero(E.p + alpha * s, nd - alpha * s, c->st);
      LB_CUDA_CHECK(cudaMemsetAsync(E_u32.p + alpha * s, 0, (nd - alpha * s) * 4, c->st));
    }
  }
  ByteWriter w;
  // ---- comm_derefs (surge.rs:136-140, subtables/mod.rs:177-184, 382-393)
  {
    SpanTimer sp(c, "Subtables.commit");
    unsigned tbits = S.kind == STRAT_LT ? 1 : (S.kind == STRAT_RANGE ? (unsigned)S.log_m : (unsigned)(S.log_m / 2));
    std::vector<uint8_t> comm = commit_u32(c, g, E_u32.p, nv_d, tbits);
    transcript.append_message("subtable_evals_commitment", std::string("begin_subtable_evals_commitment"));
    transcript.append_message("comm_poly_row_col_ops_val", std::string("poly_commitment_begin"));
    for (size_t i = 0; i < comm.size() / 32; i++) transcript.append_point_compressed("poly_commitment_share", comm.data() + 32 * i);
    transcript.append_message("comm_poly_row_col_ops_val", std::string("poly_commitment_end"));
    transcript.append_message("subtable_evals_commitment", std::string("end_subtable_evals_commitment"));
    w.vec_pts(comm);
  }
  // ---- primary sumcheck (surge.rs:142-172)
  std::vector<fr_t> r_z;
  {
    DBuf<fr_t> Wk(c, (alpha + 1) * s);  // clones of E_i + eq(r): the sumcheck binds them in place
    LB_CUDA_CHECK(cudaMemcpyAsync(Wk.p, E.p, alpha * s * sizeof(fr_t), cudaMemcpyDeviceToDevice, c->st));
    eq_evals_dev(c, r, 0, log_s, Wk.p + alpha * s);
    launch_sumcheck_claim(S, Wk.p, s, s, c->d_partial, c->d_small, c->st);  // subtables/mod.rs:186-216
    g_launches += 2;
    fr_t claimed_eval;
    c->d2h(&claimed_eval, c->d_small, sizeof(fr_t));
    transcript.append_scalar("claim_eval_scalar_product", claimed_eval);
    SumcheckProof primary = prove_arbitrary(c, S, Wk.p, s, s, transcript, r_z);
    ser_sumcheck(w, primary);
    w.fr(claimed_eval);
  }
  // ---- eval_derefs = E_i(r_z) (surge.rs:175-176) and the combined opening (177-184)
  DBuf<fr_t> eqtab(c, std::max(s, M));
  std::vector<fr_t> eval_derefs(alpha);
  {
    SpanTimer sp(c, "CombinedEval.prove");
    eq_evals_dev(c, r_z, 0, log_s, eqtab.p);
    launch_multi_dot(E.p, s, (int)alpha, eqtab.p, s, c->d_partial, c->d_small, c->st);
    g_launches += 2;
    c->d2h(eval_derefs.data(), c->d_small, alpha * sizeof(fr_t));
    w.arr_fr(eval_derefs);
    transcript.append_protocol_name("Lasso CombinedTableEvalProof");
    ser_dpl(w, prove_joint(c, g, E.p, nv_d, eval_derefs, true, "evals_ops_val", "challenge_combine_n_to_one",
                           "joint_claim_eval", r_z, transcript, tape));
  }
  // ---- memory checking (surge.rs:186-198)
  std::vector<fr_t> r_hash = transcript.challenge_vector("challenge_r_hash", 2);
  const fr_t gamma = r_hash[0], tau = r_hash[1];
  transcript.append_protocol_name("Lasso MemoryCheckingProof");
  std::vector<fr_t> rand_mem, rand_ops;
  {
    SpanTimer sp(c, "ProductLayer.prove");
    // Subtables::to_grand_products (subtables/mod.rs:133-175) + GrandProducts::new (memory_checking.rs:175-217)
    std::vector<std::unique_ptr<Circuit>> init(alpha), rd(alpha), wr(alpha), fin(alpha);
    for (size_t i = 0; i < alpha; i++) {
      size_t j = (size_t)S.memory_to_dimension_index((int)i), k = (size_t)S.memory_to_subtable_index((int)i);
      for (auto* pc : {&init[i], &fin[i]}) {
        pc->reset(new Circuit());
        (*pc)->N = M;
        (*pc)->tree.alloc(c, 2 * M);
      }
      for (auto* pc : {&rd[i], &wr[i]}) {
        pc->reset(new Circuit());
        (*pc)->N = s;
        (*pc)->tree.alloc(c, 2 * s);
      }
      launch_gp_fingerprints_mem(tables_fr.p + k * M, dense.fin(j), M, gamma, tau, init[i]->tree.p, fin[i]->tree.p, c->st);
      launch_gp_fingerprints_ops(dense.dim(j), E.p + i * s, dense.read(j), s, gamma, tau, rd[i]->tree.p, wr[i]->tree.p, c->st);
      g_launches += 2;
      build_tree(c, *init[i]);
      build_tree(c, *fin[i]);
      build_tree(c, *rd[i]);
      build_tree(c, *wr[i]);
    }
    // ProductLayerProof::prove (memory_checking.rs:673-731)
    transcript.append_protocol_name("Lasso ProductLayerProof");
    auto evaluate = [&](Circuit& ci) {  // grand_product.rs:60-65
      fr_t top[2];
      c->d2h(top, ci.layer(ci.num_layers - 1), 64);
      return fr_mul(top[0], top[1]);
    };
    std::vector<fr_t> claims_rw, claims_if;
    for (size_t i = 0; i < alpha; i++) {
      fr_t hi = evaluate(*init[i]), hr = evaluate(*rd[i]), hw = evaluate(*wr[i]), hf = evaluate(*fin[i]);
      if (!fr_eq(fr_mul(hi, hw), fr_mul(hr, hf))) throw std::runtime_error("multiset hash check failed (memory_checking.rs:689)");
      transcript.append_scalar("claim_hash_init", hi);
      transcript.append_scalar("claim_hash_read", hr);
      transcript.append_scalar("claim_hash_write", hw);
      transcript.append_scalar("claim_hash_final", hf);
      w.fr(hi);
      w.fr(hr);
      w.fr(hw);
      w.fr(hf);
      claims_rw.push_back(hr);
      claims_rw.push_back(hw);
      claims_if.push_back(hi);
      claims_if.push_back(hf);
    }
    std::vector<Circuit*> rw, inf;
    for (size_t i = 0; i < alpha; i++) {
      rw.push_back(rd[i].get());
      rw.push_back(wr[i].get());
      inf.push_back(init[i].get());
      inf.push_back(fin[i].get());
    }
    GPAProof proof_ops = prove_gpa(c, rw, claims_rw, transcript, rand_ops);
    GPAProof proof_mem = prove_gpa(c, inf, claims_if, transcript, rand_mem);
    ser_gpa(w, proof_mem);  // field order: grand_product_evals, proof_mem, proof_ops (memory_checking.rs:655-660)
    ser_gpa(w, proof_ops);
  }
  {
    // HashLayerProof::prove (memory_checking.rs:337-460)
    SpanTimer sp(c, "HashLayer.prove");
    transcript.append_protocol_name("Lasso HashLayerProof");
    std::vector<fr_t> eval_derefs2(alpha), eval_dim(C), eval_read(C), eval_final(C);
    eq_evals_dev(c, rand_ops, 0, rand_ops.size(), eqtab.p);
    launch_multi_dot(E.p, s, (int)alpha, eqtab.p, s, c->d_partial, c->d_small, c->st);
    launch_multi_dot(dense.d_l_fr.p, s, (int)(2 * C), eqtab.p, s, c->d_partial + 65536, c->d_small + 64, c->st);
    g_launches += 4;
    {
      std::vector<fr_t> tmp(64 + 2 * C);
      c->d2h(tmp.data(), c->d_small, tmp.size() * sizeof(fr_t));
      for (size_t i = 0; i < alpha; i++) eval_derefs2[i] = tmp[i];
      for (size_t i = 0; i < C; i++) {
        eval_dim[i] = tmp[64 + i];
        eval_read[i] = tmp[64 + C + i];
      }
    }
    transcript.append_protocol_name("Lasso CombinedTableEvalProof");
    DotProductProofLogBytes proof_derefs =
        prove_joint(c, g, E.p, nv_d, eval_derefs2, true, "evals_ops_val", "challenge_combine_n_to_one",
                    "joint_claim_eval", rand_ops, transcript, tape);
    eq_evals_dev(c, rand_mem, 0, rand_mem.size(), eqtab.p);
    launch_multi_dot(dense.d_m_fr.p, M, (int)C, eqtab.p, M, c->d_partial, c->d_small, c->st);
    g_launches += 2;
    c->d2h(eval_final.data(), c->d_small, C * sizeof(fr_t));
    std::vector<fr_t> evals_ops = eval_dim;
    evals_ops.insert(evals_ops.end(), eval_read.begin(), eval_read.end());
    DotProductProofLogBytes proof_ops =
        prove_joint(c, g, dense.d_l_fr.p, dense.nv_l, evals_ops, true, "claim_evals_ops", "challenge_combine_n_to_one",
                    "joint_claim_eval_ops", rand_ops, transcript, tape);
    // claim_evals_mem is appended UNPADDED and uses Math::log_2 (ceil) of C (memory_checking.rs:413-418)
    DotProductProofLogBytes proof_mem =
        prove_joint(c, g, dense.d_m_fr.p, dense.nv_m, eval_final, false, "claim_evals_mem",
                    "challenge_combine_two_to_one", "joint_claim_eval_mem", rand_mem, transcript, tape);
    // field order (memory_checking.rs:313-329)
    w.arr_fr(eval_dim);
    w.arr_fr(eval_read);
    w.arr_fr(eval_final);
    w.arr_fr(eval_derefs2);
    ser_dpl(w, proof_ops);
    ser_dpl(w, proof_mem);
    ser_dpl(w, proof_derefs);
  }
  c->sync();
  return w.b;
}

}  // namespace lb
