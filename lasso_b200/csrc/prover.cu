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
__global__ void publish_kernel(const uint32_t* src, int nwords, uint32_t* mapped, uint32_t seq) {
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) mapped[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    *((volatile uint32_t*)(mapped + 1024)) = seq;
  }
}
void Ctx::d2h_small(void* dst, const void* src, size_t bytes) {
  const uint32_t seq = ++mapped_seq;
  publish_kernel<<<1, 128, 0, st>>>((const uint32_t*)src, (int)((bytes + 3) / 4), d_mapped, seq);
  g_launches += 1;
  wait_flag(seq);
  memcpy(dst, (const void*)h_mapped, bytes);
}
void Ctx::fin_wait(const Finalize& f, fr_t* dst, int count) {
  if (f.mapped) {
    // every element arrives as one 32-byte store carrying the round's tag in bits 29..31 of its last word
    if ((size_t)count > kTaggedElems) throw std::runtime_error("round message larger than the tagged buffer");
    volatile uint32_t* w = h_mapped + kTaggedWord0;
    auto t0 = std::chrono::steady_clock::now();
    unsigned spins = 0;
    for (int v = 0; v < count; v++) {
      while ((w[8 * v + 7] >> 29) != f.tag) {
        __builtin_ia32_pause();
        if ((++spins & 0xffff) == 0) {  // surface kernel faults instead of spinning forever
          double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
          if (dt > 0.5) LB_CUDA_CHECK(cudaStreamQuery(st) == cudaErrorNotReady ? cudaSuccess : cudaStreamSynchronize(st));
          if (dt > 120.0) throw std::runtime_error("timeout waiting for a device result");
        }
      }
    }
    __sync_synchronize();
    memcpy(dst, (const void*)w, (size_t)count * sizeof(fr_t));
    for (int v = 0; v < count; v++) {
      dst[v].v[7] &= 0x1fffffffu;
      w[8 * v + 7] = 0;  // consumed: no stale tag can satisfy a later wait
    }
    return;
  }
  comm_allreduce_fr(this, d_small, count);
  d2h(dst, d_small, (size_t)count * sizeof(fr_t));
}
void Ctx::wait_points(int npoints, uint32_t* xyz) {
  volatile uint32_t* w = h_mapped + kTaggedWord0;
  const int count = 3 * npoints;
  auto t0 = std::chrono::steady_clock::now();
  unsigned spins = 0;
  for (int v = 0; v < count; v++) {
    while ((w[8 * v + 7] >> 31) == 0) {
      __builtin_ia32_pause();
      if ((++spins & 0xffff) == 0) {
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (dt > 0.5) LB_CUDA_CHECK(cudaStreamQuery(st) == cudaErrorNotReady ? cudaSuccess : cudaStreamSynchronize(st));
        if (dt > 120.0) throw std::runtime_error("timeout waiting for a device result");
      }
    }
  }
  __sync_synchronize();
  memcpy(xyz, (const void*)w, (size_t)count * 32);
  for (int v = 0; v < count; v++) {
    xyz[8 * v + 7] &= 0x7fffffffu;
    w[8 * v + 7] = 0;
  }
}
void Ctx::wait_flag(uint32_t seq) {
  volatile uint32_t* flag = h_mapped + 1024;
  auto t0 = std::chrono::steady_clock::now();
  unsigned spins = 0;
  while (*flag != seq) {
    __builtin_ia32_pause();
    if ((++spins & 0xffff) == 0) {  // surface kernel faults instead of spinning forever
      if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 0.5) {
        LB_CUDA_CHECK(cudaStreamQuery(st) == cudaErrorNotReady ? cudaSuccess : cudaStreamSynchronize(st));
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 120.0)
          throw std::runtime_error("timeout waiting for a device result");
      }
    }
  }
  __sync_synchronize();
}
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
  LB_CUDA_CHECK(cudaMemset(c->d_flag, 0, 64));
  {
    const char* nm = getenv("LASSO_B200_NO_MAPPED");
    if (!(nm && nm[0] == '1')) {
      LB_CUDA_CHECK(cudaHostAlloc((void**)&c->h_mapped, Ctx::kMappedBytes, cudaHostAllocMapped));
      memset(c->h_mapped, 0, Ctx::kMappedBytes);
      LB_CUDA_CHECK(cudaHostGetDevicePointer((void**)&c->d_mapped, c->h_mapped, 0));
    }
  }
  LB_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_aux, cudaEventDisableTiming));
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
  if (c->ev_aux) cudaEventDestroy(c->ev_aux);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  if (c->h_mapped) cudaFreeHost(c->h_mapped);
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
  g_launches += ell <= 11 ? 1 : (ell <= 22 ? 3 : 5);
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
  {
    // widest opening: R_size = 2^(nv - nv/2) generators + Q + h (dense_mlpoly.rs:301-316)
    size_t nv = std::max(g->nv_l, std::max(g->nv_m, g->nv_d));
    size_t nd = ((size_t)1 << (nv - nv / 2)) + 2;
    const char* off = getenv("LASSO_B200_NO_MULTIPLES");
    // The tables are an optimisation: if the device cannot hold them (cap, or an allocation failure on a smaller
    // or busier GPU) the prover silently keeps the bucket / 8-bit paths — outputs do not depend on it.
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const size_t bytes8 = (size_t)kMsmFullWindows * nd * 128 * sizeof(pt_niels);
    if (c->world == 1 && nd <= n_points && !(off && off[0] == '1') && bytes8 < free_b / 2) {
      g->n_direct = nd;
      g->d_multiples.alloc(c, (size_t)kMsmFullWindows * nd * 128);
      launch_build_multiples(g->d_table.p, n_points, nd, kMsmFullWindows, g->d_multiples.p, c->st);
      g_launches += 1;
      const char* cap = getenv("LASSO_B200_TABLE_GB");
      const double cap_gb = cap ? atof(cap) : 64.0;
      const size_t ncols16 = nd - 2;
      const size_t bytes16 = ncols16 * 32768 * sizeof(pt_niels);
      if ((double)bytes16 <= cap_gb * 1e9 && bytes16 < (free_b - bytes8) / 2) {
        g->n_direct16 = ncols16;
        g->d_multiples16.alloc(c, ncols16 * 32768);
        launch_build_multiples16(g->d_table.p, g->d_multiples.p, nd, ncols16, g->d_multiples16.p, c->st);
        g_launches += 1;
        g->d_centre.alloc(c, 32);
        for (size_t k = 0; ((size_t)1 << k) <= ncols16 && k < 32; k++) {  // one constant per power-of-two row length
          launch_centre_constant(g->d_multiples16.p, 1 << k, g->d_centre.p + k, c->st);
          g_launches += 1;
        }
      }
    }
  }
  c->sync();
  return g.release();
}

// ---------------------------------------------------------------------------------------------- sharding helpers
// One proof sharded over G = c->world GPUs: every array of global length n >= G is partitioned by the low
// log2(G) index bits (rank g holds X[i*G + g]); see comm.cu.  With G == 1 all of this is the identity.
static inline size_t loc(const Ctx* c, size_t n) {
  if (n % (size_t)c->world) throw std::runtime_error("array shorter than the number of GPUs");
  return n / (size_t)c->world;
}
// sum a device-side partial result over the ranks, then bring it to the host
static void reduce_to_host(Ctx* c, fr_t* d_buf, int count, fr_t* h_out) {
  comm_allreduce_fr(c, d_buf, count);
  c->d2h(h_out, d_buf, (size_t)count * sizeof(fr_t));
}
// this rank's shard of eq(r[off .. off+ell)) (eq_poly.rs:21-38): eq[i*G + g] = eq_hi[i] * eq_lo[g] where
// eq_lo is the table of the LAST log2(G) coordinates (they bind the low index bits: r[0] <-> MSB)
static void eq_evals_shard(Ctx* c, const std::vector<fr_t>& r, size_t off, size_t ell, fr_t* out) {
  const size_t lg = (size_t)c->lg_world;
  if (ell < lg) throw std::runtime_error("eq table smaller than the number of GPUs");
  eq_evals_dev(c, r, off, ell - lg, out);
  if (lg == 0) return;
  fr_t k = fr_one();
  for (size_t j = 0; j < lg; j++) {
    const fr_t& rj = r[off + ell - lg + j];
    bool bit = (c->rank >> (lg - 1 - j)) & 1;
    k = fr_mul(k, bit ? rj : fr_sub(fr_one(), rj));
  }
  launch_scale(out, out, (size_t)1 << (ell - lg), k, c->st);
  g_launches += 1;
}

// ---------------------------------------------------------------------------------------------- MSM helpers
// Row-MSMs over the generator table.  `d_scal` holds this rank's columns (ncols_loc per row, local column c'
// = global column c'*G + g + col0).  Optional replicated tail: `d_tail` = nrows x ntail scalars on the
// generators [tail_col0, tail_col0 + ntail) (the Q and h terms of a Bulletproofs round) — identical on
// every rank, so it is added once, after the cross-GPU gather.  Returns nrows compressed points.
static std::vector<uint8_t> msm_rows(Ctx* c, const Gens& g, const void* d_scal, int limbs, size_t row_stride, int nrows,
                                     int ncols_loc, int nw, const fr_t* d_tail_canon, int ntail, size_t tail_col0) {
  const int G = c->world;
  std::vector<uint8_t> out((size_t)nrows * 32);
  const bool few = nrows <= 8;
  DBuf<pt_ext> part(c, msm_partials_count(nrows, ncols_loc, nw));
  if (G == 1 && !d_tail_canon && !few) {  // the common single-GPU commit: normalise on the device
    DBuf<uint32_t> comp(c, (size_t)nrows * 8);
    launch_msm_rows(g.d_table.p, g.n_points, 1, d_scal, limbs, row_stride, nrows, ncols_loc, nw, 1, 0, part.p, nullptr,
                    comp.p, nullptr, c->st);
    g_launches += 2;
    c->d2h(out.data(), comp.p, out.size());
    return out;
  }
  // general path: raw partial points -> (gather over ranks) -> (+ tail) -> sum -> normalise
  const int nsrc = G + (d_tail_canon ? 1 : 0);
  DBuf<uint32_t> raw(c, (size_t)(nsrc + 1) * nrows * 32);
  uint32_t* mine = raw.p + (size_t)nsrc * nrows * 32;  // scratch slot for this rank's partials
  launch_msm_rows(g.d_table.p, g.n_points, 1, d_scal, limbs, row_stride, nrows, ncols_loc, nw, G, c->rank, part.p, nullptr,
                  nullptr, G == 1 ? raw.p : mine, c->st);
  g_launches += 2;
  if (G > 1) comm_allgather(c, mine, raw.p, (size_t)nrows * 128);
  if (d_tail_canon) {
    DBuf<pt_ext> tpart(c, msm_partials_count(nrows, ntail, kMsmFullWindows));
    launch_msm_rows(g.d_table.p + tail_col0, g.n_points, 1, d_tail_canon, 8, (size_t)ntail, nrows, ntail, kMsmFullWindows,
                    1, 0, tpart.p, nullptr, nullptr, raw.p + (size_t)G * nrows * 32, c->st);
    g_launches += 2;
  }
  if (few) {
    uint32_t xyzt[8 * 32];
    if (nsrc > 1) {
      launch_sum_raw_points(raw.p, nsrc, nrows, mine, nullptr, nullptr, c->st);
      g_launches += 1;
      c->d2h(xyzt, mine, (size_t)nrows * 128);
    } else {
      c->d2h(xyzt, raw.p, (size_t)nrows * 128);
    }
    // a couple of points per Bulletproofs round: invert on the host (3 us vs ~100 us on one GPU thread)
    for (int i = 0; i < nrows; i++) h64::compress_xyz(xyzt + 32 * i, out.data() + 32 * i);
    return out;
  }
  DBuf<uint32_t> comp(c, (size_t)nrows * 8);
  launch_sum_raw_points(raw.p, nsrc, nrows, nullptr, comp.p, nullptr, c->st);
  g_launches += 1;
  c->d2h(out.data(), comp.p, out.size());
  return out;
}
// rows of Montgomery Fr scalars (this rank's columns) + optional replicated Montgomery tail scalars
static std::vector<uint8_t> msm_rows_fr(Ctx* c, const Gens& g, const fr_t* d_scal_mont, int nrows, int ncols_loc,
                                        const fr_t* d_tail_mont = nullptr, int ntail = 0, size_t tail_col0 = 0) {
  DBuf<fr_t> canon(c, (size_t)nrows * ncols_loc + (size_t)nrows * ntail);
  launch_canonicalize(d_scal_mont, canon.p, (size_t)nrows * ncols_loc, c->d_flag, c->st);
  g_launches += 1;
  fr_t* tcan = nullptr;
  if (d_tail_mont) {
    tcan = canon.p + (size_t)nrows * ncols_loc;
    launch_canonicalize(d_tail_mont, tcan, (size_t)nrows * ntail, c->d_flag, c->st);
    g_launches += 1;
  }
  return msm_rows(c, g, canon.p, 8, (size_t)ncols_loc, nrows, ncols_loc, kMsmFullWindows, tcan, ntail, tail_col0);
}
// replicated tiny MSM on generators [col0, col0 + ncols): every rank computes the same point, no exchange
static std::vector<uint8_t> msm_replicated_fr(Ctx* c, const Gens& g, const fr_t* d_scal_mont, int ncols, size_t col0) {
  DBuf<fr_t> canon(c, (size_t)ncols);
  launch_canonicalize(d_scal_mont, canon.p, (size_t)ncols, c->d_flag, c->st);
  DBuf<pt_ext> part(c, msm_partials_count(1, ncols, kMsmFullWindows));
  DBuf<uint32_t> raw(c, 32);
  launch_msm_rows(g.d_table.p + col0, g.n_points, 1, canon.p, 8, (size_t)ncols, 1, ncols, kMsmFullWindows, 1, 0, part.p,
                  nullptr, nullptr, raw.p, c->st);
  g_launches += 3;
  uint32_t xyzt[32];
  c->d2h(xyzt, raw.p, 128);
  std::vector<uint8_t> out(32);
  h64::compress_xyz(xyzt, out.data());
  return out;
}

// DensePolynomial::commit (dense_mlpoly.rs:152-181) for an integer-valued polynomial of 2^nv entries viewed as
// L x R; this rank holds, for every row, the R/G columns congruent to its rank (= its low-bit shard of the array)
static std::vector<uint8_t> commit_u32(Ctx* c, const Gens& g, const uint32_t* d_vals_loc, size_t nv, unsigned max_bits) {
  size_t L = (size_t)1 << (nv / 2), R = (size_t)1 << (nv - nv / 2);
  if (R + 2 > g.n_points) throw std::runtime_error("generator stream too short for this polynomial");
  int nw = msm_windows_for_bits(max_bits);
  if (nw > 5) throw std::runtime_error("u32 MSM path: scalars wider than 32 bits");
  size_t R_loc = loc(c, R);
  if (c->world == 1 && g.d_multiples.p && R <= g.n_direct && L > 8) {
    // single GPU: rows as direct sums over the digit-multiples table (msm_kernels.cu), normalised on the device
    std::vector<uint8_t> out(L * 32);
    DBuf<pt_ext> part(c, L);
    DBuf<uint32_t> comp(c, L * 8);
    const bool wide = R <= g.n_direct16;
    launch_msm_rows_direct_u32(g.d_multiples.p, g.n_direct, wide ? g.d_multiples16.p : nullptr,
                               wide ? g.d_centre.p + (nv - nv / 2) : nullptr, d_vals_loc, R, (int)L, (int)R, nw, part.p, nullptr,
                               comp.p, nullptr, c->st);
    g_launches += 2;
    c->d2h(out.data(), comp.p, out.size());
    return out;
  }
  return msm_rows(c, g, d_vals_loc, 1, R_loc, (int)L, (int)R_loc, nw, nullptr, 0, 0);
}

// ---------------------------------------------------------------------------------------------- densify
Dense* densify(Ctx* c, const uint64_t* indices, size_t n, size_t C, size_t log_m, int* err) {
  SpanTimer sp(c, "Densify");
  *err = 0;
  if (n == 0 || C == 0 || C > 16 || log_m < 1 || log_m > 28) {
    *err = 4;
    return nullptr;
  }
  const size_t G = (size_t)c->world, gr = (size_t)c->rank;
  std::unique_ptr<Dense> d(new Dense());
  d->ctx = c;
  d->C = C;
  d->s = next_pow2(n);
  d->log_m = log_m;
  d->m = (size_t)1 << log_m;
  d->nv_l = log2_exact_or_ceil(next_pow2(2 * C * d->s));
  d->nv_m = log2_exact_or_ceil(next_pow2(C)) + log_m;
  const size_t s = d->s, m = d->m;
  if (G > 1 && (s < 2 * G || m < 2 * G)) {
    *err = 4;
    return nullptr;
  }
  d->s_loc = s / G;
  d->m_loc = m / G;
  const size_t s_loc = d->s_loc, m_loc = d->m_loc;
  const size_t nl = ((size_t)1 << d->nv_l) / G, nm = ((size_t)1 << d->nv_m) / G;  // local lengths
  {
    const char* hd = getenv("LASSO_B200_HOST_DENSIFY");
    const char* gd = getenv("LASSO_B200_GPU_DENSIFY");
    // measured on B200 (profiles/README.md): the device path wins from 2^22 lookups up (53 vs 73 ms at 2^24) and
    // whenever one proof is sharded (no replicated host scan); below that the 3 ms host scan + pinned upload wins
    const bool want_gpu = (gd && gd[0] == '1') || s >= ((size_t)1 << 22) || G > 1;
    if (densify_gpu_supported(s, log_m) && want_gpu && !(hd && hd[0] == '1')) {
      // GPU path (densify_kernels.cu): upload the raw index matrix, derive dim / read / final on the device.
      // When one proof is sharded every rank does this for the whole sequence and stores only its shard.
      d->d_l_u32.alloc(c, nl);
      d->d_m_u32.alloc(c, nm);
      d->d_l_fr.alloc(c, nl);
      d->d_m_fr.alloc(c, nm);
      // narrow usize -> u32 (and range-check, densified.rs:46) while staging into pinned memory: half the
      // PCIe bytes and a full-rate copy; a few host threads keep up with the link
      uint32_t* stage = c->stage(n * C);
      {
        const size_t total = n * C, nthreads = total >= (1u << 20) ? 4 : 1;
        std::vector<int> bad(nthreads, 0);
        auto conv = [&](size_t t) {
          size_t lo = total * t / nthreads, hi = total * (t + 1) / nthreads;
          for (size_t k = lo; k < hi; k++) {
            uint64_t a = indices[k];
            if (a >= m) {
              bad[t] = 1;
              a = 0;
            }
            stage[k] = (uint32_t)a;
          }
        };
        std::vector<std::thread> th;
        for (size_t t = 1; t < nthreads; t++) th.emplace_back(conv, t);
        conv(0);
        for (auto& t : th) t.join();
        for (int bflag : bad)
          if (bflag) {
            *err = 3;
            return nullptr;
          }
      }
      DBuf<uint32_t> d_idx(c, n * C);
      LB_CUDA_CHECK(cudaMemcpyAsync(d_idx.p, stage, n * C * sizeof(uint32_t), cudaMemcpyHostToDevice, c->st));
      const size_t B = densify_chunk(s);
      DBuf<uint32_t> d_addr(c, s), d_P(c, (s / B) * m);
      if (nl > 2 * C * s_loc) LB_CUDA_CHECK(cudaMemsetAsync(d->d_l_u32.p + 2 * C * s_loc, 0, (nl - 2 * C * s_loc) * 4, c->st));
      if (nm > C * m_loc) LB_CUDA_CHECK(cudaMemsetAsync(d->d_m_u32.p + C * m_loc, 0, (nm - C * m_loc) * 4, c->st));
      for (size_t i = 0; i < C; i++)
        g_launches += launch_densify_dim(d_idx.p, n, s, (int)C, (int)i, log_m, (int)G, (int)gr, d_addr.p, d_P.p,
                                         d->d_l_u32.p + i * s_loc, d->d_l_u32.p + (C + i) * s_loc,
                                         d->d_m_u32.p + i * m_loc, c->st);
      launch_from_u32(d->d_l_u32.p, d->d_l_fr.p, nl, c->st);  // DensePolynomial::from_usize + merge
      launch_from_u32(d->d_m_u32.p, d->d_m_fr.p, nm, c->st);
      g_launches += 2;
      c->sync();  // the pinned staging buffer is reused by the next call
      return d.release();
    }
  }
  // host path (memories larger than 2^16 cells): pinned staging, reused across calls: no per-call page faults, and the upload runs at full PCIe rate
  uint32_t* l_host = c->stage(nl + nm + (G > 1 ? (2 * s + m) * C : 0));
  uint32_t* m_host = l_host + nl;
  uint32_t* full = m_host + nm;  // G > 1: whole-sequence scratch (every rank runs the full scan, keeps its shard)
  if (nl > 2 * C * s_loc) memset(l_host + 2 * C * s_loc, 0, (nl - 2 * C * s_loc) * sizeof(uint32_t));
  if (nm > C * m_loc) memset(m_host + C * m_loc, 0, (nm - C * m_loc) * sizeof(uint32_t));
  // densified.rs:33-56: per dimension, pad with address 0 and run the (inherently sequential) timestamp
  // counters; dimensions are independent, so one host thread each.
  std::vector<int> bad(C, 0);
  auto work = [&](size_t i) {
    uint32_t* dim = G == 1 ? l_host + i * s : full + i * (2 * s + m);
    uint32_t* rd = G == 1 ? l_host + (C + i) * s : dim + s;
    uint32_t* fin = G == 1 ? m_host + i * m : dim + 2 * s;
    memset(fin, 0, m * sizeof(uint32_t));
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
    if (G > 1) {  // keep the low-bit shard
      uint32_t* ld = l_host + i * s_loc;
      uint32_t* lr = l_host + (C + i) * s_loc;
      uint32_t* lf = m_host + i * m_loc;
      for (size_t k = 0; k < s_loc; k++) {
        ld[k] = dim[k * G + gr];
        lr[k] = rd[k * G + gr];
      }
      for (size_t k = 0; k < m_loc; k++) lf[k] = fin[k * G + gr];
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
// sumcheck.rs:149-260 over device polynomials W_k = base + k*stride (k <= alpha, the last is eq); `len_loc` is
// this rank's length.  Sharded rounds: local eval -> sum over ranks -> host; local bind.  When one element
// per rank is left the G-element remainders are all-gathered and the last log2(G) rounds run replicated.
static SumcheckProof prove_arbitrary(Ctx* c, const Strategy& S, fr_t* base, size_t stride, size_t len_loc,
                                     Transcript& transcript, std::vector<fr_t>& r) {
  SpanTimer sp(c, "Sumcheck.prove");
  SumcheckProof proof;
  r.clear();
  const int npts = S.sumcheck_poly_degree() + 1, npolys = S.num_memories() + 1;
  std::vector<fr_t> evals(npts);
  DBuf<fr_t> tail;
  bool sharded = c->world > 1;
  size_t len = len_loc;
  for (;;) {
    if (sharded && len == 1) {  // hand over to the replicated tail
      tail.alloc(c, (size_t)npolys * c->world);
      comm_gather_heads(c, nullptr, base, stride, npolys, tail.p);
      base = tail.p;
      stride = (size_t)c->world;
      len = (size_t)c->world;
      sharded = false;
    }
    if (len <= 1) break;
    size_t half = len / 2;
    if (sharded) {
      Finalize f = c->fin_begin();
      launch_sumcheck_eval_arbitrary(S, base, stride, half, f, c->st);
      c->fin_wait(f, evals.data(), npts);
    } else {  // replicated tail of a sharded proof, or a single GPU: no cross-rank sum
      Finalize f = c->fin_begin();
      launch_sumcheck_eval_arbitrary(S, base, stride, half, f, c->st);
      if (f.mapped)
        c->fin_wait(f, evals.data(), npts);
      else
        c->d2h(evals.data(), c->d_small, (size_t)npts * sizeof(fr_t));
    }
    g_launches += 1;
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
// Sharded: layers with N/2^k >= G are held as low-bit shards (local length N/(2^k G)); the layer of global
// length G is all-gathered and the few layers above it are kept replicated on every rank.
struct Circuit {
  DBuf<fr_t> tree;   // local shards: layer 0 at 0 (N/G elements), layer 1 after it, ...
  DBuf<fr_t> rtree;  // replicated top: layer k_rep (G elements), k_rep + 1, ...   (empty when G == 1)
  size_t N = 0, num_layers = 0;
  int G = 1;
  size_t k_rep = 0;  // first replicated layer: N >> k_rep == G
  bool layer_is_sharded(size_t k) const { return G == 1 || (N >> k) >= 2 * (size_t)G; }
  size_t layer_len_global(size_t k) const { return N >> k; }
  fr_t* layer_local(size_t k) const {  // valid for (N >> k) >= G
    size_t off = 0, len = N / G;
    for (size_t i = 0; i < k; i++) {
      off += len;
      len /= 2;
    }
    return tree.p + off;
  }
  fr_t* layer_rep(size_t k) const {  // valid for k >= k_rep (G > 1)
    size_t off = 0, len = (size_t)G;
    for (size_t i = k_rep; i < k; i++) {
      off += len;
      len /= 2;
    }
    return rtree.p + off;
  }
};
static void circuit_alloc(Ctx* c, Circuit& ci, size_t N) {
  ci.N = N;
  ci.G = c->world;
  ci.num_layers = log2_exact_or_ceil(N);
  ci.tree.alloc(c, 2 * (N / ci.G));
  if (ci.G > 1) {
    ci.k_rep = ci.num_layers - (size_t)c->lg_world;
    ci.rtree.alloc(c, 2 * (size_t)ci.G);
  }
}
static void build_tree(Ctx* c, Circuit& ci) {  // grand_product.rs:38-58 (layer 0 already filled)
  const size_t last_local = ci.G == 1 ? ci.num_layers - 1 : ci.k_rep;
  for (size_t k = 0; k < last_local; k++) {
    launch_product_layer(ci.layer_local(k), ci.layer_local(k + 1), ci.layer_len_global(k + 1) / ci.G, c->st);
    g_launches += 1;
  }
  if (ci.G > 1) {
    comm_gather_heads(c, nullptr, ci.layer_local(ci.k_rep), 0, 1, ci.layer_rep(ci.k_rep));
    for (size_t k = ci.k_rep; k + 1 < ci.num_layers; k++) {
      launch_product_layer(ci.layer_rep(k), ci.layer_rep(k + 1), ci.layer_len_global(k + 1), c->st);
      g_launches += 1;
    }
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
  const int ncirc = (int)circuits.size(), G = c->world;
  const size_t num_layers = circuits[0]->num_layers;
  // pointer tables: slot L (< num_layers) = the arrays of layer L, slot num_layers = the replicated tail arrays;
  // all of them are uploaded once, up front (no per-layer copy + sync)
  const size_t nslots = num_layers + 1;
  DBuf<fr_t*> d_ptrs(c, nslots * 4 * ncirc);
  const size_t eq_cap = std::max<size_t>(circuits[0]->N / 2 / G, (size_t)G);
  DBuf<fr_t> eqbuf(c, eq_cap), eqbuf2(c, std::max<size_t>(eq_cap / 2, 1));
  DBuf<fr_t> tail(c, (size_t)(2 * ncirc + 1) * G);  // replicated remainders of A_k, B_k, C (G elements each)
  std::vector<fr_t> rand;
  if (ncirc > 32) throw std::runtime_error("more than 32 circuits in one batched grand product");
  std::vector<fr_t> ev(3), fin((size_t)2 * ncirc);
  // per slot: [A_0..A_{n-1} | B_0..B_{n-1} | A_0,B_0,A_1,B_1,...]
  std::vector<fr_t*> table(nslots * 4 * ncirc);
  auto slot_A = [&](size_t slot) { return d_ptrs.p + slot * 4 * ncirc; };
  auto slot_B = [&](size_t slot) { return d_ptrs.p + slot * 4 * ncirc + ncirc; };
  auto slot_AB = [&](size_t slot) { return d_ptrs.p + slot * 4 * ncirc + 2 * ncirc; };
  auto layer_cur = [&](size_t layer_id, bool& replicated_layer) {
    const size_t len_g = circuits[0]->layer_len_global(layer_id);
    replicated_layer = G > 1 && !circuits[0]->layer_is_sharded(layer_id);
    return replicated_layer ? len_g / 2 : len_g / 2 / (size_t)G;  // |A| = |B| = |C| on this rank
  };
  for (size_t slot = 0; slot < nslots; slot++) {
    for (int k = 0; k < ncirc; k++) {
      fr_t *pa, *pb;
      if (slot == num_layers) {
        pa = tail.p + (size_t)(2 * k) * G;
        pb = tail.p + (size_t)(2 * k + 1) * G;
      } else {
        bool rep;
        size_t cur0 = layer_cur(slot, rep);
        pa = rep ? circuits[k]->layer_rep(slot) : circuits[k]->layer_local(slot);
        pb = pa + cur0;
      }
      table[slot * 4 * ncirc + k] = pa;
      table[slot * 4 * ncirc + ncirc + k] = pb;
      table[slot * 4 * ncirc + 2 * ncirc + 2 * k] = pa;
      table[slot * 4 * ncirc + 2 * ncirc + 2 * k + 1] = pb;
    }
  }
  LB_CUDA_CHECK(cudaMemcpyAsync(d_ptrs.p, table.data(), table.size() * sizeof(fr_t*), cudaMemcpyHostToDevice, c->st));
  c->sync();
  for (size_t layer_id = num_layers; layer_id-- > 0;) {
    bool replicated_layer;
    size_t cur = layer_cur(layer_id, replicated_layer);
    bool sharded = G > 1 && !replicated_layer;
    fr_t* const* dA = slot_A(layer_id);
    fr_t* const* dB = slot_B(layer_id);
    fr_t* const* dAB = slot_AB(layer_id);
    // poly_C = eq(rand), grand_product.rs:122
    if (sharded)
      eq_evals_shard(c, rand, 0, rand.size(), eqbuf.p);
    else
      eq_evals_dev(c, rand, 0, rand.size(), eqbuf.p);
    std::vector<fr_t> coeff_vec = transcript.challenge_vector("rand_coeffs_next_layer", ncirc);
    fr_t e = fr_zero();
    for (int k = 0; k < ncirc; k++) e = fr_add(e, fr_mul(claims_to_verify[k], coeff_vec[k]));
    // The kernels fold the batching coefficients in (poly_kernels.cu): the first bind of the layer stores
    // coeff_k * A_k, a round message is the 3 combined values of sumcheck.rs:95-97.
    CubicCoeffs cf;
    for (int k = 0; k < ncirc; k++) cf.v[k] = coeff_vec[k];
    bool stored_scaled = false;
    std::vector<fr_t> inv_coeff;  // computed while the first kernel of the layer runs
    LayerProof lp;
    std::vector<fr_t> rand_prod;
    fr_t* Ccur = eqbuf.p;
    fr_t* Cnext = eqbuf2.p;
    bool have_evals = false, heads_published = false;
    Finalize fz = c->fin_begin();
    for (;;) {
      if (sharded && cur == 1) {  // all-gather the G-element remainders; the tail rounds run replicated
        comm_gather_heads(c, dAB, nullptr, 0, 2 * ncirc, tail.p);
        comm_gather_heads(c, nullptr, Ccur, 0, 1, tail.p + (size_t)2 * ncirc * G);
        dA = slot_A(num_layers);
        dB = slot_B(num_layers);
        dAB = slot_AB(num_layers);
        Ccur = tail.p + (size_t)2 * ncirc * G;
        Cnext = eqbuf2.p;
        cur = (size_t)G;
        sharded = false;
        have_evals = false;
      }
      if (cur <= 1) break;
      if (!have_evals) {  // first round of a phase; later rounds come out of the fused bind+eval kernel
        fz = c->fin_begin();
        launch_sumcheck_eval_cubic_comb(dA, dB, Ccur, ncirc, cur / 2, cf, stored_scaled ? 0 : 1, fz, c->st);
        g_launches += 1;
      }
      if (inv_coeff.empty()) {  // 1 / coeff_k by Montgomery's trick, overlapping the kernel just launched
        inv_coeff.resize(ncirc);
        std::vector<fr_t> pre(ncirc);
        fr_t acc = fr_one();
        for (int k = 0; k < ncirc; k++) {
          pre[k] = acc;
          acc = fr_mul(acc, coeff_vec[k]);
        }
        if (fr_eq(acc, fr_zero())) throw std::runtime_error("zero batching coefficient");
        fr_t ainv = fr_inv(acc);
        for (int k = ncirc; k-- > 0;) {
          inv_coeff[k] = fr_mul(ainv, pre[k]);
          ainv = fr_mul(ainv, coeff_vec[k]);
        }
      }
      size_t half = cur / 2;
      auto tp0 = std::chrono::steady_clock::now();
      if (sharded || fz.mapped)
        c->fin_wait(fz, ev.data(), 3);
      else
        c->d2h(ev.data(), c->d_small, ev.size() * sizeof(fr_t));
      auto tp1 = std::chrono::steady_clock::now();
      const fr_t c0 = ev[0], c2 = ev[1], c3 = ev[2];  // already combined over the circuits (sumcheck.rs:95-97)
      std::vector<fr_t> evals = {c0, fr_sub(e, c0), c2, c3};  // eval(1) = e - eval(0), sumcheck.rs:99-104
      std::vector<fr_t> coeffs = unipoly_from_evals(evals);
      unipoly_append(coeffs, transcript);
      fr_t r_j = transcript.challenge_scalar("challenge_nextround");
      rand_prod.push_back(r_j);
      auto tp2 = std::chrono::steady_clock::now();
      if (half > 1) {
        // bind with r_j and evaluate the next round in one pass (sumcheck.rs:116-120 + 63-89)
        fz = c->fin_begin();
        launch_sumcheck_bind_eval_cubic_comb(dA, dB, Ccur, Cnext, ncirc, half, r_j, cf, stored_scaled ? 0 : 1, fz, c->st);
        stored_scaled = true;
        g_launches += 1;
        std::swap(Ccur, Cnext);
        have_evals = true;
      } else if (fz.mapped) {
        // last round: bind the 2*ncirc heads and publish them (the layer's claims); eq is not needed any more
        fz = c->fin_begin();
        launch_bind_heads(dAB, 2 * ncirc, r_j, fz, c->st);
        g_launches += 1;
        have_evals = false;
        heads_published = true;
      } else {
        launch_bind_top_ptrs(dAB, 2 * ncirc, half, r_j, c->st);
        launch_bind_top(Ccur, 0, 1, half, r_j, c->st);
        g_launches += 2;
        have_evals = false;
      }
      auto tp3 = std::chrono::steady_clock::now();
      if (c->span_sync) {  // where a grand-product round goes: waiting for the device, host glue, launch call
        c->spans["GPA.round wait"] += std::chrono::duration<double, std::milli>(tp1 - tp0).count();
        c->spans["GPA.round host"] += std::chrono::duration<double, std::milli>(tp2 - tp1).count();
        c->spans["GPA.round launch"] += std::chrono::duration<double, std::milli>(tp3 - tp2).count();
      }
      e = unipoly_evaluate(coeffs, r_j);
      lp.proof.push_back(unipoly_compress(coeffs));
      cur = half;
    }
    // claims_prod = (A_k[0], B_k[0]): published by the last round's kernel, or packed on the device + one transfer
    if (heads_published) {
      c->fin_wait(fz, fin.data(), 2 * ncirc);
    } else {
      pack_heads(c, dAB, nullptr, 0, 2 * ncirc, c->d_small + 1024);
      c->d2h(fin.data(), c->d_small + 1024, fin.size() * sizeof(fr_t));
    }
    for (int k = 0; k < ncirc; k++) {  // the left arrays carry coeff_k once a bind has stored them
      lp.claims_prod_left.push_back(stored_scaled ? fr_mul(fin[2 * k], inv_coeff[k]) : fin[2 * k]);
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

// dst (nrows x 2, row-major) <- [ip[0], blind_L ; ip[1], blind_R]
__global__ void set_tail_kernel(fr_t* dst0, fr_t* dst1, const fr_t* ip, fr_t blind_L, fr_t blind_R) {
  if (threadIdx.x || blockIdx.x) return;
  dst0[0] = ip[0];  // c_L on Q
  dst0[1] = blind_L;  // on H
  dst1[0] = ip[1];
  dst1[1] = blind_R;
}
__global__ void set_elems_kernel(fr_t* dst, fr_t a, fr_t b) {
  if (threadIdx.x || blockIdx.x) return;
  dst[0] = a;
  dst[1] = b;
}

// PolyEvalProof::prove (dense_mlpoly.rs:301-359) -> DotProductProofLog::prove (dot_product.rs:166-249)
// -> BulletReductionProof::prove (bullet.rs:40-154).  Z: this rank's shard of a polynomial of 2^nv elements,
// i.e. for every one of the L rows the R/G columns congruent to the rank.
static DotProductProofLogBytes prove_poly_eval(Ctx* c, const Gens& g, const fr_t* Z, size_t nv,
                                               const std::vector<fr_t>& r, const fr_t& Zr, Transcript& transcript,
                                               RandomTape& tape) {
  SpanTimer sp(c, "DensePolyEval.prove");
  transcript.append_protocol_name("polynomial evaluation proof");
  if (r.size() != nv) throw std::runtime_error("PolyEvalProof: r.len() != num_vars");
  const int G = c->world, gr = c->rank;
  const size_t lv = nv / 2, rv = nv - nv / 2, L_size = (size_t)1 << lv, n = (size_t)1 << rv;  // n = R_size
  if (n + 2 > g.n_points) throw std::runtime_error("generator stream too short");
  if (n < 2 * (size_t)G) throw std::runtime_error("opening narrower than 2 x #GPUs");
  const size_t lg_n = rv, n_loc = n / G;
  // L, R = factored eq evals (eq_poly.rs:44-52); LZ = L . Z (dense_mlpoly.rs:183-207)
  std::unique_ptr<SpanTimer> sp1(new SpanTimer(c, "PE.1 eq+bound"));
  DBuf<fr_t> Lvec(c, L_size), a(c, n_loc), b(c, n_loc), arep(c, 2 * (size_t)G);
  eq_evals_dev(c, r, 0, lv, Lvec.p);          // rows are not sharded: L is replicated
  eq_evals_shard(c, r, lv, rv, b.p);          // a_vec of the dot product proof = R (this rank's columns)
  if ((size_t)bound_max_chunks() * n_loc > c->partial_elems) throw std::runtime_error("bound scratch too small");
  launch_bound(Z, Lvec.p, L_size, n_loc, c->d_partial, a.p, c->st);  // x_vec = LZ
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
  // single-GPU pipeline below: needs the multiples table of the generators 0 .. n+1
  const bool fast = G == 1 && c->h_mapped != nullptr && n * 32 <= c->h_pin_bytes && g.d_multiples.p && n + 2 <= g.n_direct;
  DBuf<fr_t> two(c, 4);
  if (!fast) {
    // Cx = batch_commit(x_vec, blind_x = 0) ; Cy = y*Q + 0*h
    std::vector<uint8_t> Cx = msm_rows_fr(c, g, a.p, 1, (int)n_loc);
    transcript.append_point_compressed("Cx", Cx.data());
    set_elems_kernel<<<1, 32, 0, c->st>>>(two.p, Zr, fr_zero());
    g_launches += 1;
    std::vector<uint8_t> Cy = msm_replicated_fr(c, g, two.p, 2, n);
    transcript.append_point_compressed("Cy", Cy.data());
    {  // append_scalars(b"a", a_vec): canonical bytes straight from the device (all ranks need the whole vector)
      DBuf<fr_t> canon(c, n_loc), all(c, G > 1 ? n : 0);
      launch_canonicalize(b.p, canon.p, n_loc, c->d_flag, c->st);
      g_launches += 1;
      std::vector<uint8_t> bytes(n * 32);
      if (G == 1) {
        c->d2h(bytes.data(), canon.p, bytes.size());
      } else {
        comm_allgather(c, canon.p, all.p, n_loc * 32);
        std::vector<uint8_t> tmp(n * 32);
        c->d2h(tmp.data(), all.p, tmp.size());
        for (int q = 0; q < G; q++)  // rank q's local column j' is global column j'*G + q
          for (size_t j = 0; j < n_loc; j++) memcpy(&bytes[(j * G + q) * 32], &tmp[((size_t)q * n_loc + j) * 32], 32);
      }
      transcript.append_scalars_bytes("a", bytes.data(), n);
    }
  }
  // ---- BulletReductionProof::prove with unfolded generators (see file header)
  if (!fast) sp1.reset(new SpanTimer(c, "PE.3 bullet rounds"));
  fr_t blind_fin = fr_zero();  // blind_Gamma = blind_x + blind_y = 0
  const size_t ncols_main = G == 1 ? n + 2 : n_loc;  // single GPU: Q and h ride along as columns n, n+1
  DBuf<fr_t> W0(c, n), W1(c, n), sLR(c, 2 * ncols_main), tailsc(c, 4);
  fr_t* W = W0.p;   // weights of the unfolded generators: replicated (indexed by the HIGH column bits)
  fr_t* Wn = W1.p;
  set_elems_kernel<<<1, 32, 0, c->st>>>(W, fr_one(), fr_zero());
  g_launches += 1;
  fr_t* sL = sLR.p;
  fr_t* sR = sLR.p + ncols_main;
  fr_t* av = a.p;  // current a / b vectors: sharded while m >= 2G, replicated afterwards
  fr_t* bv = b.p;
  DBuf<fr_t> a_alt, b_alt;
  if (fast) {
    // Single GPU.  Per message ONE scalar kernel + the two MSM kernels, the finish kernel publishing straight to
    // mapped host memory; the host part of a message (compression, Fiat-Shamir) overlaps the device work of the
    // next one wherever the transcript allows it.
    //   (Cx, Cy): rows (x_vec, 0, 0) and (0.., y, 0) of one two-row MSM        (dot_product.rs:192-197)
    //   round k : fold with u_{k-1}, weights, L/R scalars, c_L, c_R -> two-row MSM   (bullet.rs:73-134)
    auto read_two_points = [&](uint8_t* comp64) {
      uint32_t xyz[48];
      c->wait_points(2, xyz);
      h64::compress_xyz_pair(xyz, xyz + 24, comp64, comp64 + 32);  // one Fq inversion for both points
    };
    a_alt.alloc(c, n);
    b_alt.alloc(c, n);
    fr_t *an = a_alt.p, *bn = b_alt.p;
    DBuf<pt_ext> part(c, 2 * (size_t)msm_direct_chunks((int)(n + 2), 1));
    DBuf<fr_t> canon(c, n);
    DBuf<uint32_t> cols(c, 2 * (n / 2 + 2));
    // two short rows over the multiples table; len terms per row, generator index per term in cols (or identity)
    // heavy = rows that carry the terms: both in a round (L, R), one for (Cx, Cy) — Cy is a single term
    auto two_row_msm = [&](const uint32_t* d_cols, size_t len, int heavy) {
      launch_msm_direct(g.d_multiples.p, g.n_direct, (const uint32_t*)sLR.p, d_cols, 2, (int)len, heavy, part.p, nullptr,
                        c->d_mapped + Ctx::kTaggedWord0, c->st);
      g_launches += 2;
    };
    launch_two_row_scalars(av, 0, fr_one(), fr_zero(), fr_zero(), Zr, fr_zero(), n, sLR.p, c->st);
    two_row_msm(nullptr, n + 2, 1);
    // a_vec of the transcript = canonical bytes of b; the copy is waited for only when it is appended
    launch_canonicalize(bv, canon.p, n, c->d_flag, c->st);
    LB_CUDA_CHECK(cudaMemcpyAsync(c->h_pin, canon.p, n * 32, cudaMemcpyDeviceToHost, c->st));
    LB_CUDA_CHECK(cudaEventRecord(c->ev_aux, c->st));
    g_launches += 2;
    uint8_t CxCy[64];
    read_two_points(CxCy);
    fr_t u = fr_one(), u_inv = fr_one();
    int fold = 0;
    size_t m = n;  // vector length entering the round (after the fold with the previous challenge)
    auto launch_round = [&](size_t round) {
      launch_bullet_round(av, bv, W, an, bn, Wn, n, m, fold, u, u_inv, v1[round], v2[round], sLR.p, cols.p, c->d_partial,
                          c->d_flag + 4, c->st);
      g_launches += 1;
      if (fold) {
        std::swap(av, an);
        std::swap(bv, bn);
        std::swap(W, Wn);
      }
      two_row_msm(cols.p, n / 2 + 2, 2);
    };
    if (m != 1) launch_round(0);  // round 0 needs no challenge: it runs while the host absorbs Cx, Cy, a
    transcript.append_point_compressed("Cx", CxCy);
    transcript.append_point_compressed("Cy", CxCy + 32);
    LB_CUDA_CHECK(cudaEventSynchronize(c->ev_aux));
    transcript.append_scalars_bytes("a", c->h_pin, n);
    sp1.reset(new SpanTimer(c, "PE.3 bullet rounds"));
    for (size_t round = 0; m != 1; round++) {
      uint8_t LR[64];
      read_two_points(LR);
      transcript.append_point_compressed("L", LR);
      transcript.append_point_compressed("R", LR + 32);
      u = transcript.challenge_scalar("u");
      u_inv = fr_inv(u);
      fold = 1;
      m /= 2;
      if (m != 1) launch_round(round + 1);
      blind_fin = fr_add(blind_fin, fr_add(fr_mul(fr_mul(v1[round], u), u), fr_mul(fr_mul(v2[round], u_inv), u_inv)));
      out.L_vec.insert(out.L_vec.end(), LR, LR + 32);
      out.R_vec.insert(out.R_vec.end(), LR + 32, LR + 64);
    }
    if (fold) {  // the last challenge: a, b -> one element each, weights -> n (bullet.rs:127-134)
      launch_fold_ab(av, bv, 1, u, u_inv, c->st);
      launch_expand_weights(W, Wn, n / 2, u, u_inv, c->st);
      g_launches += 2;
      std::swap(W, Wn);
    }
  } else {
  bool sharded = G > 1;
  size_t m = n, nw_count = 1;  // current (global) vector length, number of weights
  for (size_t round = 0; m != 1; round++) {
    if (sharded && m == (size_t)G) {  // one element per rank left: gather, finish replicated
      comm_gather_heads(c, nullptr, av, 0, 1, arep.p);
      comm_gather_heads(c, nullptr, bv, 0, 1, arep.p + G);
      av = arep.p;
      bv = arep.p + G;
      sharded = false;
    }
    const size_t h = m / 2;
    const size_t h_arr = sharded ? h / G : h;  // half length of the arrays this rank holds
    launch_cross_inner_products(av, bv, h_arr, c->d_partial, c->d_small, c->st);  // c_L, c_R (bullet.rs:78-79)
    g_launches += 2;
    if (sharded) comm_allreduce_fr(c, c->d_small, 2);
    launch_bullet_scalars(av, W, n_loc, m, G, gr, (G > 1 && !sharded) ? 1 : 0, sL, sR, c->st);
    g_launches += 1;
    std::vector<uint8_t> LR;
    if (G == 1) {
      set_tail_kernel<<<1, 32, 0, c->st>>>(sL + n, sR + n, c->d_small, v1[round], v2[round]);
      g_launches += 1;
      LR = msm_rows_fr(c, g, sLR.p, 2, (int)(n + 2));
    } else {
      set_tail_kernel<<<1, 32, 0, c->st>>>(tailsc.p, tailsc.p + 2, c->d_small, v1[round], v2[round]);
      g_launches += 1;
      LR = msm_rows_fr(c, g, sLR.p, 2, (int)n_loc, tailsc.p, 2, n);
    }
    transcript.append_point_compressed("L", LR.data());
    transcript.append_point_compressed("R", LR.data() + 32);
    fr_t u = transcript.challenge_scalar("u");
    fr_t u_inv = fr_inv(u);
    launch_fold_ab(av, bv, h_arr, u, u_inv, c->st);  // bullet.rs:127-130 (scalars only; G stays unfolded)
    launch_expand_weights(W, Wn, nw_count, u, u_inv, c->st);
    g_launches += 2;
    std::swap(W, Wn);
    nw_count *= 2;
    blind_fin = fr_add(blind_fin, fr_add(fr_mul(fr_mul(v1[round], u), u), fr_mul(fr_mul(v2[round], u_inv), u_inv)));
    out.L_vec.insert(out.L_vec.end(), LR.begin(), LR.begin() + 32);
    out.R_vec.insert(out.R_vec.end(), LR.begin() + 32, LR.begin() + 64);
    m = h;
  }
  }
  sp1.reset(new SpanTimer(c, "PE.4 delta,beta"));
  fr_t ab[2];
  if (fast) {
    // delta = d * g_hat + r_delta * h with g_hat = sum_j W[j] G_j (dot_product.rs:219-227) and
    // beta = d * Q + r_beta * h (dot_product.rs:229-230) as the two rows of one MSM
    launch_two_row_scalars(W, 1, d, fr_zero(), r_delta, d, r_beta, n, sLR.p, c->st);
    DBuf<pt_ext> part(c, 2 * (size_t)msm_direct_chunks((int)(n + 2), 1));
    launch_msm_direct(g.d_multiples.p, g.n_direct, (const uint32_t*)sLR.p, nullptr, 2, (int)(n + 2), 1, part.p, nullptr,
                      c->d_mapped + Ctx::kTaggedWord0, c->st);
    LB_CUDA_CHECK(cudaMemcpyAsync(c->h_pin, av, 32, cudaMemcpyDeviceToHost, c->st));
    LB_CUDA_CHECK(cudaMemcpyAsync(c->h_pin + 32, bv, 32, cudaMemcpyDeviceToHost, c->st));
    g_launches += 3;
    uint32_t xyz[48];
    c->wait_points(2, xyz);
    h64::compress_xyz_pair(xyz, xyz + 24, out.delta, out.beta);
    c->sync();
    memcpy(ab, c->h_pin, 64);
    transcript.append_point_compressed("delta", out.delta);
    transcript.append_point_compressed("beta", out.beta);
  } else {
    LB_CUDA_CHECK(cudaMemcpyAsync(c->h_pin, av, 32, cudaMemcpyDeviceToHost, c->st));
    LB_CUDA_CHECK(cudaMemcpyAsync(c->h_pin + 32, bv, 32, cudaMemcpyDeviceToHost, c->st));
    c->sync();
    memcpy(ab, c->h_pin, 64);
    // delta = d * g_hat + r_delta * h with g_hat = sum_j W[j] G_j  (dot_product.rs:219-227)
    std::vector<uint8_t> delta;
    if (G == 1) {
      launch_scale(W, sL, n, d, c->st);
      set_elems_kernel<<<1, 32, 0, c->st>>>(sL + n, fr_zero(), r_delta);
      g_launches += 2;
      delta = msm_rows_fr(c, g, sL, 1, (int)(n + 2));
    } else {
      launch_scale_strided(W, sL, n_loc, (size_t)G, (size_t)gr, d, c->st);
      set_elems_kernel<<<1, 32, 0, c->st>>>(tailsc.p, fr_zero(), r_delta);
      g_launches += 2;
      delta = msm_rows_fr(c, g, sL, 1, (int)n_loc, tailsc.p, 2, n);
    }
    memcpy(out.delta, delta.data(), 32);
    transcript.append_point_compressed("delta", out.delta);
    // beta = d * Q + r_beta * h  (dot_product.rs:229-230)
    set_elems_kernel<<<1, 32, 0, c->st>>>(two.p, d, r_beta);
    g_launches += 1;
    std::vector<uint8_t> beta = msm_replicated_fr(c, g, two.p, 2, n);
    memcpy(out.beta, beta.data(), 32);
    transcript.append_point_compressed("beta", out.beta);
  }
  fr_t x_hat = ab[0], a_hat = ab[1], rhat_Gamma = blind_fin;
  fr_t y_hat = fr_mul(x_hat, a_hat);
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
  const int G = c->world, gr = c->rank;
  const size_t s = dense.s, C = dense.C, M = dense.m, alpha = (size_t)S.num_memories();
  const size_t s_loc = dense.s_loc, M_loc = dense.m_loc;
  const size_t log_s = log2_exact_or_ceil(s);
  if ((size_t)S.C != C || (size_t)S.log_m != dense.log_m) throw std::runtime_error("strategy does not match the densified representation");
  if (g.nv_d != log2_exact_or_ceil(next_pow2(alpha * s)) || g.nv_l != dense.nv_l || g.nv_m != dense.nv_m)
    throw std::runtime_error("generators were built for different (c, s, num_memories, log_m)");
  transcript.append_protocol_name("Lasso SparsePolynomialEvaluationProof");

  // ---- Subtables::new (subtables/mod.rs:116-129): materialise (replicated, 2-6 MiB), gather, merge
  const size_t nv_d = g.nv_d, nd_loc = ((size_t)1 << nv_d) / G;
  const int nsub = S.num_subtables();
  DBuf<fr_t> tables_fr(c, (size_t)nsub * M);
  DBuf<uint32_t> tables_u32(c, (size_t)nsub * M);
  DBuf<fr_t> E(c, nd_loc);          // combined_poly = E_0 | .. | E_{alpha-1} | 0-pad (this rank's shard)
  DBuf<uint32_t> E_u32(c, nd_loc);  // same values as integers for the small-scalar commit
  {
    SpanTimer sp(c, "Subtables.new");
    launch_materialize_subtables(S, tables_fr.p, tables_u32.p, c->st);
    launch_gather_lookup_polys(S, tables_fr.p, tables_u32.p, dense.nz(), s_loc, E.p, s_loc, E_u32.p, c->st);
    g_launches += 2;
    if (nd_loc > alpha * s_loc) {
      launch_fill_zero(E.p + alpha * s_loc, nd_loc - alpha * s_loc, c->st);
      LB_CUDA_CHECK(cudaMemsetAsync(E_u32.p + alpha * s_loc, 0, (nd_loc - alpha * s_loc) * 4, c->st));
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
    DBuf<fr_t> Wk(c, (alpha + 1) * s_loc);  // clones of E_i + eq(r): the sumcheck binds them in place
    LB_CUDA_CHECK(cudaMemcpyAsync(Wk.p, E.p, alpha * s_loc * sizeof(fr_t), cudaMemcpyDeviceToDevice, c->st));
    eq_evals_shard(c, r, 0, log_s, Wk.p + alpha * s_loc);
    launch_sumcheck_claim(S, Wk.p, s_loc, s_loc, c->d_partial, c->d_small, c->st);  // subtables/mod.rs:186-216
    g_launches += 2;
    fr_t claimed_eval;
    reduce_to_host(c, c->d_small, 1, &claimed_eval);
    transcript.append_scalar("claim_eval_scalar_product", claimed_eval);
    SumcheckProof primary = prove_arbitrary(c, S, Wk.p, s_loc, s_loc, transcript, r_z);
    ser_sumcheck(w, primary);
    w.fr(claimed_eval);
  }
  // ---- eval_derefs = E_i(r_z) (surge.rs:175-176) and the combined opening (177-184)
  DBuf<fr_t> eqtab(c, std::max(s_loc, M_loc));
  std::vector<fr_t> eval_derefs(alpha);
  {
    SpanTimer sp(c, "CombinedEval.prove");
    eq_evals_shard(c, r_z, 0, log_s, eqtab.p);
    launch_multi_dot(E.p, s_loc, (int)alpha, eqtab.p, s_loc, c->d_partial, c->d_small, c->st);
    g_launches += 2;
    reduce_to_host(c, c->d_small, (int)alpha, eval_derefs.data());
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
        circuit_alloc(c, **pc, M);
      }
      for (auto* pc : {&rd[i], &wr[i]}) {
        pc->reset(new Circuit());
        circuit_alloc(c, **pc, s);
      }
      launch_gp_fingerprints_mem(tables_fr.p + k * M, dense.fin(j), M_loc, G, gr, gamma, tau, init[i]->tree.p,
                                 fin[i]->tree.p, c->st);
      launch_gp_fingerprints_ops(dense.dim(j), E.p + i * s_loc, dense.read(j), s_loc, gamma, tau, rd[i]->tree.p,
                                 wr[i]->tree.p, c->st);
      g_launches += 2;
    }
    // all trees of a size at once + the top layers straight to the host (single GPU); else tree by tree
    const bool batched = G == 1 && c->h_mapped && 2 * alpha <= 32;
    std::vector<fr_t> tops(8 * alpha);
    if (batched) {
      TreePtrs tm, to;
      for (size_t i = 0; i < alpha; i++) {
        tm.p[2 * i] = init[i]->tree.p;
        tm.p[2 * i + 1] = fin[i]->tree.p;
        to.p[2 * i] = rd[i]->tree.p;
        to.p[2 * i + 1] = wr[i]->tree.p;
      }
      Finalize f = c->fin_begin();
      launch_product_trees(tm, (int)(2 * alpha), M, 0, f, c->st);
      launch_product_trees(to, (int)(2 * alpha), s, (int)(2 * alpha), f, c->st);
      g_launches += product_trees_launches(M) + product_trees_launches(s);
      c->fin_wait(f, tops.data(), (int)(8 * alpha));
    } else {
      for (size_t i = 0; i < alpha; i++) {
        build_tree(c, *init[i]);
        build_tree(c, *fin[i]);
        build_tree(c, *rd[i]);
        build_tree(c, *wr[i]);
      }
    }
    // ProductLayerProof::prove (memory_checking.rs:673-731)
    transcript.append_protocol_name("Lasso ProductLayerProof");
    auto evaluate = [&](Circuit& ci, size_t slot) {  // grand_product.rs:60-65 (the top layer is replicated when G > 1)
      if (batched) return fr_mul(tops[2 * slot], tops[2 * slot + 1]);
      fr_t top[2];
      c->d2h(top, G == 1 ? ci.layer_local(ci.num_layers - 1) : ci.layer_rep(ci.num_layers - 1), 64);
      return fr_mul(top[0], top[1]);
    };
    std::vector<fr_t> claims_rw, claims_if;
    for (size_t i = 0; i < alpha; i++) {
      fr_t hi = evaluate(*init[i], 2 * i), hr = evaluate(*rd[i], 2 * alpha + 2 * i),
           hw = evaluate(*wr[i], 2 * alpha + 2 * i + 1), hf = evaluate(*fin[i], 2 * i + 1);
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
    eq_evals_shard(c, rand_ops, 0, rand_ops.size(), eqtab.p);
    launch_multi_dot(E.p, s_loc, (int)alpha, eqtab.p, s_loc, c->d_partial, c->d_small, c->st);
    launch_multi_dot(dense.d_l_fr.p, s_loc, (int)(2 * C), eqtab.p, s_loc, c->d_partial + 65536, c->d_small + 64, c->st);
    g_launches += 4;
    {
      std::vector<fr_t> tmp(64 + 2 * C);
      reduce_to_host(c, c->d_small, (int)tmp.size(), tmp.data());
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
    eq_evals_shard(c, rand_mem, 0, rand_mem.size(), eqtab.p);
    launch_multi_dot(dense.d_m_fr.p, M_loc, (int)C, eqtab.p, M_loc, c->d_partial, c->d_small, c->st);
    g_launches += 2;
    reduce_to_host(c, c->d_small, (int)C, eval_final.data());
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
