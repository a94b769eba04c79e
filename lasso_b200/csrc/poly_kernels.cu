// lasso_b200 — CUDA kernels (sm_100a) for the multilinear-polynomial side of the Lasso prover
// hot path: bind (K1), sumcheck round evaluation (K2, K3), eq evals (K4), subtable
// materialisation + gather (K5), and the supporting reductions (K7).  SURVEY.md §2.2.
//
// All of these are streaming integer kernels over 32-byte field elements: one element per
// thread per 256-bit load (a warp covers 1 KiB contiguous), grid sized as a multiple of the
// 148 SMs, grid-stride loops, warp-shuffle + shared-memory reductions for partial sums.
// No tensor cores: this is 256-bit modular integer arithmetic, not a dense contraction.
#include "kernels.cuh"

namespace lb {

static constexpr int kThreads = 256;
static constexpr int kBlocksPerSM = 4;
static constexpr int kMaxBlocks = kNumSMs * kBlocksPerSM;  // 592
// bind_top launch shape, measured with tools/bind_sweep.py on 5 x 2^22 elements (GB/s of the 96 B per output):
// CTAs/SM x outputs per thread: 4x1 5340, 4x2 5514, 5x1 5496, 5x2 5661, 6x1 5655, 6x2 5582
static constexpr int kBindBlocksPerSM = 5, kBindIlp = 2;

static inline int grid_for(size_t n, int threads = kThreads, int max_blocks = kMaxBlocks) {
  size_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > (size_t)max_blocks) b = max_blocks;
  return (int)b;
}
int sumcheck_max_blocks() { return kMaxBlocks; }

// ------------------------------------------------------------------------------------ K1
// dense_mlpoly.rs:209-216 — algorithmic traffic 96 B per output element, 1 modmul.
__global__ void __launch_bounds__(kThreads) bind_top_kernel(fr_t* base, size_t stride, size_t half, fr_t r) {
  fr_t* Z = base + (size_t)blockIdx.y * stride;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
    fr_t lo = ld_fr_stream(Z + i), hi = ld_fr_stream(Z + half + i);
    st_fr(Z + i, fr_add(lo, fr_mul(r, fr_sub(hi, lo))));
  }
}
__global__ void __launch_bounds__(kThreads) bind_top_ptrs_kernel(fr_t* const* ptrs, size_t half, fr_t r) {
  fr_t* Z = ptrs[blockIdx.y];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
    fr_t lo = ld_fr_stream(Z + i), hi = ld_fr_stream(Z + half + i);
    st_fr(Z + i, fr_add(lo, fr_mul(r, fr_sub(hi, lo))));
  }
}
// dense_mlpoly.rs:218-225
__global__ void __launch_bounds__(kThreads) bind_bot_kernel(const fr_t* Z, fr_t* out, size_t half, fr_t r) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
    fr_t lo = ld_fr(Z + 2 * i), hi = ld_fr(Z + 2 * i + 1);
    st_fr(out + i, fr_add(lo, fr_mul(r, fr_sub(hi, lo))));
  }
}
// two outputs per thread and iteration: four independent 32-byte loads in flight before the first multiplication
__global__ void __launch_bounds__(kThreads) bind_top2_kernel(fr_t* base, size_t stride, size_t half, fr_t r) {
  fr_t* Z = base + (size_t)blockIdx.y * stride;
  const size_t step = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + step < half; i += 2 * step) {
    const fr_t lo0 = ld_fr_stream(Z + i), hi0 = ld_fr_stream(Z + half + i);
    const fr_t lo1 = ld_fr_stream(Z + i + step), hi1 = ld_fr_stream(Z + half + i + step);
    st_fr(Z + i, fr_add(lo0, fr_mul(r, fr_sub(hi0, lo0))));
    st_fr(Z + i + step, fr_add(lo1, fr_mul(r, fr_sub(hi1, lo1))));
  }
  if (i < half) {
    const fr_t lo = ld_fr_stream(Z + i), hi = ld_fr_stream(Z + half + i);
    st_fr(Z + i, fr_add(lo, fr_mul(r, fr_sub(hi, lo))));
  }
}
// experiment knobs (tools/bind_sweep.py): resident CTAs per SM the grid is sized for, outputs per thread and iteration
static int bind_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
void launch_bind_top(fr_t* base, size_t stride, int npolys, size_t half, const fr_t& r, cudaStream_t st) {
  if (half == 0 || npolys == 0) return;
  static const int bps = bind_env("LASSO_B200_BIND_BLOCKS", kBindBlocksPerSM), ilp = bind_env("LASSO_B200_BIND_ILP", kBindIlp);
  int per = kNumSMs * bps / npolys;
  if (per < kNumSMs / 4) per = kNumSMs / 4;
  dim3 grid(grid_for(half, kThreads, per), npolys);
  if (ilp == 2)
    bind_top2_kernel<<<grid, kThreads, 0, st>>>(base, stride, half, r);
  else
    bind_top_kernel<<<grid, kThreads, 0, st>>>(base, stride, half, r);
  LB_LAUNCH_CHECK();
}
void launch_bind_top_ptrs(fr_t* const* d_ptrs, int npolys, size_t half, const fr_t& r, cudaStream_t st) {
  if (half == 0 || npolys == 0) return;
  int per = kMaxBlocks / npolys;
  if (per < kNumSMs / 4) per = kNumSMs / 4;
  dim3 grid(grid_for(half, kThreads, per), npolys);
  bind_top_ptrs_kernel<<<grid, kThreads, 0, st>>>(d_ptrs, half, r);
  LB_LAUNCH_CHECK();
}
void launch_bind_bot(const fr_t* Z, fr_t* out, size_t half, const fr_t& r, cudaStream_t st) {
  if (half == 0) return;
  bind_bot_kernel<<<grid_for(half), kThreads, 0, st>>>(Z, out, half, r);
  LB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------ K4
// eq_poly.rs:21-38.  evals[i] = prod_j (bit_{l-1-j}(i) ? r_j : 1 - r_j), r[0] <-> MSB.
// Small tables by the reference's doubling recurrence inside one CTA; big tables as the outer
// product T_hi (x) T_lo: one modmul and one 32-byte write per output.
__global__ void __launch_bounds__(1024) eq_small_kernel(FrVec r, int r_off, int ell, fr_t* out) {
  // out has 2^ell entries, ell <= 12
  if (threadIdx.x == 0) out[0] = fr_one();
  __syncthreads();
  int size = 1;
  for (int j = 0; j < ell; j++) {
    fr_t rj = r.v[r_off + j];
    fr_t old[2];
    int cnt = 0;
    for (int i = threadIdx.x; i < size; i += blockDim.x) old[cnt++] = out[i];
    __syncthreads();
    cnt = 0;
    for (int i = threadIdx.x; i < size; i += blockDim.x) {
      fr_t s = old[cnt++];
      fr_t hi = fr_mul(s, rj);
      out[2 * i + 1] = hi;
      out[2 * i] = fr_sub(s, hi);
    }
    __syncthreads();
    size *= 2;
  }
}
__global__ void __launch_bounds__(kThreads) eq_outer_kernel(const fr_t* t_hi, const fr_t* t_lo, int ell_lo,
                                                            size_t n, fr_t* out) {
  size_t mask = ((size_t)1 << ell_lo) - 1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    fr_t a = t_hi[i >> ell_lo], b = t_lo[i & mask];
    st_fr(out + i, fr_mul(a, b));
  }
}
void launch_eq_evals(const FrVec& r, int ell, fr_t* out, fr_t* scratch, cudaStream_t st) {
  if (ell <= 11) {
    eq_small_kernel<<<1, 1024, 0, st>>>(r, 0, ell, out);
    LB_LAUNCH_CHECK();
    return;
  }
  int ell_lo = ell / 2 > 11 ? 11 : ell / 2;
  int ell_hi = ell - ell_lo;
  if (ell_hi > 11) {
    // > 2^22 entries: the high table (2^ell_hi <= 2^17 entries) is itself an outer product.  Build it in the
    // tail of `out`, stage it in scratch (+4096) because the final pass overwrites that tail, then expand.
    fr_t* hi_tab = out + (((size_t)1 << ell) - ((size_t)1 << ell_hi));
    launch_eq_evals(r, ell_hi, hi_tab, scratch, st);
    eq_small_kernel<<<1, 1024, 0, st>>>(r, ell_hi, ell_lo, scratch);
    LB_LAUNCH_CHECK();
    cudaMemcpyAsync(scratch + 4096, hi_tab, sizeof(fr_t) << ell_hi, cudaMemcpyDeviceToDevice, st);
    size_t n = (size_t)1 << ell;
    eq_outer_kernel<<<grid_for(n), kThreads, 0, st>>>(scratch + 4096, scratch, ell_lo, n, out);
    LB_LAUNCH_CHECK();
    return;
  }
  eq_small_kernel<<<1, 1024, 0, st>>>(r, 0, ell_hi, scratch);
  LB_LAUNCH_CHECK();
  eq_small_kernel<<<1, 1024, 0, st>>>(r, ell_hi, ell_lo, scratch + 4096);
  LB_LAUNCH_CHECK();
  size_t n = (size_t)1 << ell;
  eq_outer_kernel<<<grid_for(n), kThreads, 0, st>>>(scratch, scratch + 4096, ell_lo, n, out);
  LB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------ partial-sum reduce
// partial: [nv][nblocks]; out[v] = sum_b partial[v][b].  One CTA per value.
__global__ void __launch_bounds__(kThreads) reduce_partials_kernel(const fr_t* partial, int nblocks, fr_t* out) {
  __shared__ fr_t scratch[kThreads / 32];
  const fr_t* p = partial + (size_t)blockIdx.x * nblocks;
  fr_t acc[1] = {fr_zero()};
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) acc[0] = fr_add(acc[0], p[i]);
  block_sum_fr<1>(acc, scratch);
  if (threadIdx.x == 0) out[blockIdx.x] = acc[0];
}

// ------------------------------------------------------------------------------------ K2
// sumcheck.rs:179-237 for the strategies whose g is linear in the E_k:
//   g(E, eq) = (sum_k 2^(k*inc) E_k) * eq      (and.rs:45-53, or.rs, xor.rs, range_check.rs:78-86)
// degree 2 -> evaluation points t = 0, 1, 2 with P(t) = lo + t (hi - lo) built incrementally.
// The weighted sum is a Horner chain of shifts (fr_mul_pow2: ~35 instructions against ~250 for a Montgomery
// multiplication), which leaves 3 multiplications per index pair and makes the big rounds HBM-bound.
// Reads 2 * 32 B per polynomial per index pair (64 B/pair/poly algorithmic).
__global__ void __launch_bounds__(kThreads)
    sc_eval_linear_kernel(const fr_t* base, size_t stride, int alpha, size_t half, int inc, Finalize fin) {
  __shared__ fr_t scratch[3 * kThreads / 32];
  fr_t acc[3] = {fr_zero(), fr_zero(), fr_zero()};
  const fr_t* eq = base + (size_t)alpha * stride;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
    const fr_t* P = base + (size_t)(alpha - 1) * stride;
    fr_t c0 = ld_fr_stream(P + i), c1 = ld_fr_stream(P + half + i);
    for (int k = alpha - 2; k >= 0; k--) {
      P = base + (size_t)k * stride;
      c0 = fr_add(fr_mul_pow2(c0, inc), ld_fr_stream(P + i));
      c1 = fr_add(fr_mul_pow2(c1, inc), ld_fr_stream(P + half + i));
    }
    fr_t q0 = ld_fr_stream(eq + i), q1 = ld_fr_stream(eq + half + i);
    acc[0] = fr_add(acc[0], fr_mul(c0, q0));
    acc[1] = fr_add(acc[1], fr_mul(c1, q1));
    fr_t c2 = fr_sub(fr_dbl(c1), c0), q2 = fr_sub(fr_dbl(q1), q0);
    acc[2] = fr_add(acc[2], fr_mul(c2, q2));
  }
  block_sum_fr<3>(acc, scratch);
  finalize_block<3>(fin, acc, 0, blockIdx.x, gridDim.x, 3, gridDim.x);
}

// The bind of round j-1 (sumcheck.rs:247-253, dense_mlpoly.rs:209-216) and the evaluation of round j in ONE pass
// over the polynomials: thread i owns the four elements i, i+q, i+2q, i+3q of every polynomial (q = a quarter of
// the length before the bind), folds them to the two elements i, i+q of the bound polynomial, stores those in place
// and feeds them to round j's sums.  192 B per poly per i instead of 96 + 96 (bind) + 64 + 64 (evaluation).
template <int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
    sc_bind_eval_linear_kernel(fr_t* base, size_t stride, int alpha, size_t q, fr_t r, int inc, Finalize fin) {
  __shared__ fr_t scratch[3 * kThreads / 32];
  fr_t acc[3] = {fr_zero(), fr_zero(), fr_zero()};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < q; i += (size_t)gridDim.x * blockDim.x) {
    fr_t c0 = fr_zero(), c1 = fr_zero();
    for (int k = alpha - 1; k >= 0; k--) {
      fr_t* P = base + (size_t)k * stride;
      const fr_t a0 = ld_fr_stream(P + i), a1 = ld_fr_stream(P + q + i);
      const fr_t a2 = ld_fr_stream(P + 2 * q + i), a3 = ld_fr_stream(P + 3 * q + i);
      const fr_t n0 = fr_add(a0, fr_mul(r, fr_sub(a2, a0))), n1 = fr_add(a1, fr_mul(r, fr_sub(a3, a1)));
      st_fr(P + i, n0);
      st_fr(P + q + i, n1);
      c0 = fr_add(fr_mul_pow2(c0, inc), n0);
      c1 = fr_add(fr_mul_pow2(c1, inc), n1);
    }
    fr_t* E = base + (size_t)alpha * stride;
    const fr_t e0 = ld_fr_stream(E + i), e1 = ld_fr_stream(E + q + i);
    const fr_t e2 = ld_fr_stream(E + 2 * q + i), e3 = ld_fr_stream(E + 3 * q + i);
    const fr_t q0 = fr_add(e0, fr_mul(r, fr_sub(e2, e0))), q1 = fr_add(e1, fr_mul(r, fr_sub(e3, e1)));
    st_fr(E + i, q0);
    st_fr(E + q + i, q1);
    acc[0] = fr_add(acc[0], fr_mul(c0, q0));
    acc[1] = fr_add(acc[1], fr_mul(c1, q1));
    const fr_t c2 = fr_sub(fr_dbl(c1), c0), q2 = fr_sub(fr_dbl(q1), q0);
    acc[2] = fr_add(acc[2], fr_mul(c2, q2));
  }
  block_sum_fr<3>(acc, scratch);
  finalize_block<3>(fin, acc, 0, blockIdx.x, gridDim.x, 3, gridDim.x);
}

// LT strategy (lt.rs:60-69): g = sum_i LT_i prod_{j<i} EQ_j, memories ordered LT_0, EQ_0, LT_1, ...
// Evaluated by Horner from the last pair, h_t <- LT_k(t) + EQ_k(t) * h_t, for all C+2 points t at once
// so only two polynomials' values are live at a time.
template <int C>
__global__ void __launch_bounds__(128)
    sc_eval_lt_kernel(const fr_t* base, size_t stride, size_t half, Finalize fin) {
  constexpr int NP = C + 2;  // degree C+1
  __shared__ fr_t scratch[NP * 128 / 32];
  fr_t acc[NP];
#pragma unroll
  for (int t = 0; t < NP; t++) acc[t] = fr_zero();
  const fr_t* eq = base + (size_t)(2 * C) * stride;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
    fr_t h[NP];
#pragma unroll
    for (int t = 0; t < NP; t++) h[t] = fr_zero();
#pragma unroll 1
    for (int k = C - 1; k >= 0; k--) {
      const fr_t* PL = base + (size_t)(2 * k) * stride;
      const fr_t* PE = base + (size_t)(2 * k + 1) * stride;
      fr_t l0 = ld_fr(PL + i), l1 = ld_fr(PL + half + i);
      fr_t e0 = ld_fr(PE + i), e1 = ld_fr(PE + half + i);
      fr_t dl = fr_sub(l1, l0), de = fr_sub(e1, e0);
      fr_t cl = l0, ce = e0;
#pragma unroll
      for (int t = 0; t < NP; t++) {
        h[t] = fr_add(cl, fr_mul(ce, h[t]));
        cl = fr_add(cl, dl);
        ce = fr_add(ce, de);
      }
    }
    fr_t q0 = ld_fr(eq + i), q1 = ld_fr(eq + half + i);
    fr_t dq = fr_sub(q1, q0), cq = q0;
#pragma unroll
    for (int t = 0; t < NP; t++) {
      acc[t] = fr_add(acc[t], fr_mul(h[t], cq));
      cq = fr_add(cq, dq);
    }
  }
  block_sum_fr<NP>(acc, scratch);
  finalize_block<NP>(fin, acc, 0, blockIdx.x, gridDim.x, NP, gridDim.x);
}

// The same evaluation with TWO lanes per index pair, each forming half of the C+2 evaluation points: the live state per
// thread halves (5 running products + 5 sums instead of 10 + 10 for C = 8: 255 registers and 8 warps per SM in the
// kernel above), the two lanes of a pair load the same addresses (a warp still reads 512 contiguous bytes per array).
// Sums over the lanes of equal parity by xor-shuffles, then as block_sum_fr.
template <int C>
__global__ void __launch_bounds__(128, 3)
    sc_eval_lt2_kernel(const fr_t* base, size_t stride, size_t half, Finalize fin) {
  constexpr int NP = C + 2, HP = (NP + 1) / 2;
  __shared__ fr_t scratch[NP * 128 / 32];
  const int part = threadIdx.x & 1, t0 = part * HP, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  fr_t acc[HP];
#pragma unroll
  for (int t = 0; t < HP; t++) acc[t] = fr_zero();
  const fr_t* eq = base + (size_t)(2 * C) * stride;
  const size_t pairs_per_step = ((size_t)gridDim.x * blockDim.x) >> 1;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1; i < half; i += pairs_per_step) {
    fr_t h[HP];
#pragma unroll
    for (int t = 0; t < HP; t++) h[t] = fr_zero();
#pragma unroll 1
    for (int k = C - 1; k >= 0; k--) {
      const fr_t* PL = base + (size_t)(2 * k) * stride;
      const fr_t* PE = base + (size_t)(2 * k + 1) * stride;
      const fr_t l0 = ld_fr(PL + i), l1 = ld_fr(PL + half + i), e0 = ld_fr(PE + i), e1 = ld_fr(PE + half + i);
      const fr_t dl = fr_sub(l1, l0), de = fr_sub(e1, e0);
      fr_t cl = l0, ce = e0;
      if (part) {  // start at t = HP: HP additions (cheaper than a multiplication by the constant)
#pragma unroll
        for (int j = 0; j < HP; j++) {
          cl = fr_add(cl, dl);
          ce = fr_add(ce, de);
        }
      }
#pragma unroll
      for (int t = 0; t < HP; t++) {
        h[t] = fr_add(cl, fr_mul(ce, h[t]));
        cl = fr_add(cl, dl);
        ce = fr_add(ce, de);
      }
    }
    const fr_t q0 = ld_fr(eq + i), q1 = ld_fr(eq + half + i), dq = fr_sub(q1, q0);
    fr_t cq = q0;
    if (part) {
#pragma unroll
      for (int j = 0; j < HP; j++) cq = fr_add(cq, dq);
    }
#pragma unroll
    for (int t = 0; t < HP; t++) {
      acc[t] = fr_add(acc[t], fr_mul(h[t], cq));
      cq = fr_add(cq, dq);
    }
  }
  // lanes of equal parity: xor-shuffles with strides 16 .. 2; lane 0 / 1 then hold the warp's sums of the two halves
#pragma unroll
  for (int t = 0; t < HP; t++) {
    fr_t a = acc[t];
#pragma unroll
    for (int d = 16; d >= 2; d >>= 1) {
      fr_t o;
#pragma unroll
      for (int l = 0; l < 8; l++) o.v[l] = __shfl_xor_sync(0xffffffffu, a.v[l], d);
      a = fr_add(a, o);
    }
    if (lane < 2 && t0 + t < NP) scratch[(t0 + t) * nwarps + warp] = a;
  }
  __syncthreads();
  fr_t vals[NP];
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NP; k++) {
      const fr_t x = lane < nwarps ? scratch[k * nwarps + lane] : fr_zero();
      vals[k] = warp_sum_fr(x);
    }
  }
  __syncthreads();
  finalize_block<NP>(fin, vals, 0, blockIdx.x, gridDim.x, NP, gridDim.x);
}

// Any other C (the reference is generic in C, lt.rs:13-14): the C+2 evaluation points are processed TB at a time so
// the live state stays in registers whatever C is; every pass re-reads the polynomials (a fallback, not a hot path).
// tv.v[t] = F::from(t).
template <int TB>
__global__ void __launch_bounds__(128)
    sc_eval_lt_generic_kernel(const fr_t* base, size_t stride, size_t half, int C, FrVec tv, Finalize fin) {
  __shared__ fr_t scratch[TB * 128 / 32];
  __shared__ fr_t res[32];
  __shared__ int s_last;
  const int NP = C + 2;
  const fr_t* eq = base + (size_t)(2 * C) * stride;
  for (int t0 = 0; t0 < NP; t0 += TB) {
    fr_t acc[TB];
#pragma unroll
    for (int t = 0; t < TB; t++) acc[t] = fr_zero();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
      fr_t h[TB];
#pragma unroll
      for (int t = 0; t < TB; t++) h[t] = fr_zero();
#pragma unroll 1
      for (int k = C - 1; k >= 0; k--) {
        const fr_t* PL = base + (size_t)(2 * k) * stride;
        const fr_t* PE = base + (size_t)(2 * k + 1) * stride;
        const fr_t l0 = ld_fr(PL + i), l1 = ld_fr(PL + half + i), e0 = ld_fr(PE + i), e1 = ld_fr(PE + half + i);
        const fr_t dl = fr_sub(l1, l0), de = fr_sub(e1, e0);
        fr_t cl = fr_add(l0, fr_mul(tv.v[t0], dl)), ce = fr_add(e0, fr_mul(tv.v[t0], de));
#pragma unroll
        for (int t = 0; t < TB; t++) {
          h[t] = fr_add(cl, fr_mul(ce, h[t]));
          cl = fr_add(cl, dl);
          ce = fr_add(ce, de);
        }
      }
      const fr_t q0 = ld_fr(eq + i), q1 = ld_fr(eq + half + i), dq = fr_sub(q1, q0);
      fr_t cq = fr_add(q0, fr_mul(tv.v[t0], dq));
#pragma unroll
      for (int t = 0; t < TB; t++) {
        acc[t] = fr_add(acc[t], fr_mul(h[t], cq));
        cq = fr_add(cq, dq);
      }
    }
    block_sum_fr<TB>(acc, scratch);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int t = 0; t < TB; t++)
        if (t0 + t < NP) res[t0 + t] = acc[t];
    }
    __syncthreads();
  }
  if (gridDim.x == 1) {
    if ((int)threadIdx.x < NP) finalize_publish(fin, threadIdx.x, res[threadIdx.x]);
    return;
  }
  if ((int)threadIdx.x < NP) fin.partial[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = res[threadIdx.x];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(fin.counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  finalize_last_stage(fin, gridDim.x, NP);
}

// combine_lookups weights are F::from(1u64 << (i * inc)): inc = log2 of the chunk size (and.rs:45-53, range_check.rs:78-86)
static int linear_inc(const Strategy& S) { return S.kind == STRAT_RANGE ? S.log_m : S.log_m / 2; }

template <int C>
static void launch_lt(const fr_t* base, size_t stride, size_t half, const Finalize& fin, int blocks, cudaStream_t st) {
  sc_eval_lt_kernel<C><<<blocks, 128, 0, st>>>(base, stride, half, fin);
  LB_LAUNCH_CHECK();
}

void launch_sumcheck_eval_arbitrary(const Strategy& S, const fr_t* base, size_t stride, size_t half, const Finalize& fin,
                                    cudaStream_t st) {
  int blocks;
  if (S.kind == STRAT_LT) {
    blocks = grid_for(half, 128, kMaxBlocks);
    switch (S.C) {
      case 1: launch_lt<1>(base, stride, half, fin, blocks, st); break;
      case 2: launch_lt<2>(base, stride, half, fin, blocks, st); break;
      case 3: launch_lt<3>(base, stride, half, fin, blocks, st); break;
      case 4: launch_lt<4>(base, stride, half, fin, blocks, st); break;
      case 8:
        if (half >= 4096) {  // throughput-bound rounds: two lanes per pair (3 CTAs per SM instead of 2, no spills)
          sc_eval_lt2_kernel<8><<<grid_for(2 * half, 128, kNumSMs * 6), 128, 0, st>>>(base, stride, half, fin);
          LB_LAUNCH_CHECK();
        } else {
          launch_lt<8>(base, stride, half, fin, blocks, st);
        }
        break;
      default: {
        FrVec tv;
        for (int t = 0; t < 32; t++) tv.v[t] = fr_from_u64((uint64_t)t);
        sc_eval_lt_generic_kernel<6><<<blocks, 128, 0, st>>>(base, stride, half, S.C, tv, fin);
        LB_LAUNCH_CHECK();
      }
    }
  } else {
    blocks = grid_for(half);
    sc_eval_linear_kernel<<<blocks, kThreads, 0, st>>>(base, stride, S.num_memories(), half, linear_inc(S), fin);
    LB_LAUNCH_CHECK();
  }
}
// bind with r (length 4q -> 2q) then evaluate the round over the bound polynomials, one launch; only the strategies
// with a linear g have a fused kernel (false: the caller binds and evaluates separately)
// Measured (tools/ab_primary.py, XOR C=4): at 2^24 lookups the fused rounds take 5 % off Sumcheck.prove; below
// q = 2^15 a round is a latency chain and the longer per-thread chain of the fused kernel costs ~0.7 us more than
// the two short launches it replaces — those rounds stay unfused (min_q = 0: that default; 1: always fuse).
bool launch_sumcheck_bind_eval_arbitrary(const Strategy& S, fr_t* base, size_t stride, size_t q, const fr_t& r,
                                         const Finalize& fin, size_t min_q, cudaStream_t st) {
  static const size_t dflt_min_q = (size_t)bind_env("LASSO_B200_FUSED_MIN_Q", 1 << 15);
  if (S.kind == STRAT_LT || q == 0 || q < (min_q ? min_q : dflt_min_q)) return false;
  // resident CTAs per SM the kernel is compiled for: 2 (128 registers) or 3 (80 registers, a few spilled words);
  // the grid is one wave of those
  static const int minb = bind_env("LASSO_B200_FUSED_MINB", 2);
  const int alpha = S.num_memories(), inc = linear_inc(S);
  if (minb == 3)
    sc_bind_eval_linear_kernel<3><<<grid_for(q, kThreads, kNumSMs * 3), kThreads, 0, st>>>(base, stride, alpha, q, r, inc, fin);
  else
    sc_bind_eval_linear_kernel<2><<<grid_for(q, kThreads, kNumSMs * 2), kThreads, 0, st>>>(base, stride, alpha, q, r, inc, fin);
  LB_LAUNCH_CHECK();
  return true;
}

// subtables/mod.rs:186-216: sum_k eq[k] * g(E_1[k], ..., E_alpha[k]) over the whole hypercube
__global__ void __launch_bounds__(kThreads)
    claim_linear_kernel(const fr_t* base, size_t stride, int alpha, size_t n, int inc, fr_t* partial) {
  __shared__ fr_t scratch[kThreads / 32];
  fr_t acc[1] = {fr_zero()};
  const fr_t* eq = base + (size_t)alpha * stride;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    fr_t c = ld_fr_stream(base + (size_t)(alpha - 1) * stride + i);
    for (int k = alpha - 2; k >= 0; k--) c = fr_add(fr_mul_pow2(c, inc), ld_fr_stream(base + (size_t)k * stride + i));
    acc[0] = fr_add(acc[0], fr_mul(c, ld_fr_stream(eq + i)));
  }
  block_sum_fr<1>(acc, scratch);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc[0];
}
__global__ void __launch_bounds__(kThreads)
    claim_lt_kernel(const fr_t* base, size_t stride, int C, size_t n, fr_t* partial) {
  __shared__ fr_t scratch[kThreads / 32];
  fr_t acc[1] = {fr_zero()};
  const fr_t* eq = base + (size_t)(2 * C) * stride;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    fr_t h = fr_zero();
    for (int k = C - 1; k >= 0; k--) {
      fr_t l = ld_fr(base + (size_t)(2 * k) * stride + i), e = ld_fr(base + (size_t)(2 * k + 1) * stride + i);
      h = fr_add(l, fr_mul(e, h));
    }
    acc[0] = fr_add(acc[0], fr_mul(h, ld_fr(eq + i)));
  }
  block_sum_fr<1>(acc, scratch);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc[0];
}
void launch_sumcheck_claim(const Strategy& S, const fr_t* base, size_t stride, size_t n, fr_t* partial, fr_t* out,
                           cudaStream_t st) {
  int blocks = grid_for(n);
  if (S.kind == STRAT_LT)
    claim_lt_kernel<<<blocks, kThreads, 0, st>>>(base, stride, S.C, n, partial);
  else
    claim_linear_kernel<<<blocks, kThreads, 0, st>>>(base, stride, S.num_memories(), n, linear_inc(S), partial);
  LB_LAUNCH_CHECK();
  reduce_partials_kernel<<<1, kThreads, 0, st>>>(partial, blocks, out);
  LB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------ K3
// sumcheck.rs:49-93: per circuit (e0, e2, e3) = sum_i A B C at t = 0, 2, 3.
__global__ void __launch_bounds__(kThreads)
    sc_eval_cubic_kernel(fr_t* const* A, fr_t* const* B, const fr_t* Ceq, size_t half, Finalize fin) {
  __shared__ fr_t scratch[3 * kThreads / 32];
  const fr_t* a = A[blockIdx.y];
  const fr_t* b = B[blockIdx.y];
  fr_t acc[3] = {fr_zero(), fr_zero(), fr_zero()};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
    fr_t a0 = ld_fr_stream(a + i), a1 = ld_fr_stream(a + half + i);
    fr_t b0 = ld_fr_stream(b + i), b1 = ld_fr_stream(b + half + i);
    fr_t c0 = ld_fr(Ceq + i), c1 = ld_fr(Ceq + half + i);
    acc[0] = fr_add(acc[0], fr_mul(fr_mul(a0, b0), c0));
    fr_t da = fr_sub(a1, a0), db = fr_sub(b1, b0), dc = fr_sub(c1, c0);
    fr_t a2 = fr_add(a1, da), b2 = fr_add(b1, db), c2 = fr_add(c1, dc);
    acc[1] = fr_add(acc[1], fr_mul(fr_mul(a2, b2), c2));
    fr_t a3 = fr_add(a2, da), b3 = fr_add(b2, db), c3 = fr_add(c2, dc);
    acc[2] = fr_add(acc[2], fr_mul(fr_mul(a3, b3), c3));
  }
  block_sum_fr<3>(acc, scratch);
  // value index = circuit*3 + t
  finalize_block<3>(fin, acc, blockIdx.y * 3, blockIdx.x, gridDim.x, 3 * gridDim.y, gridDim.x * gridDim.y);
}
// ---- the same rounds with the batching coefficients folded in (what the prover runs) --------------------------
// prove_cubic_batched only ever uses  sum_k coeff_k * (e0, e2, e3)_k  (sumcheck.rs:95-104), and everything in a
// round is linear in A_k.  So the FIRST bind of a layer stores coeff_k * A_k, every later round works on the
// scaled arrays, and a round evaluates  sum_i C_i(t) * sum_k A'_k,i(t) B_k,i(t):  per element pair 7
// multiplications per circuit (4 binds + 3 products) + 5 shared (2 eq binds + 3 times C(t)) instead of 12 per
// circuit; the round message is 3 elements instead of 3 per circuit.  The layer's claims A_k(r) are recovered on the
// host with the inverse coefficients (one batch inversion per layer, off the critical path).
//   scale != 0: the arrays A_k in memory are still unscaled; multiply by coeff_k on the fly (and store the scaled
//   value when binding).  Circuits are strided over blockIdx.y so that small rounds still fill the machine.
__global__ void __launch_bounds__(kThreads)
    sc_eval_cubic_comb_kernel(fr_t* const* A, fr_t* const* B, const fr_t* Ceq, size_t half, int ncirc, CubicCoeffs cf,
                              int scale, Finalize fin) {
  __shared__ fr_t scratch[3 * kThreads / 32];
  fr_t acc[3] = {fr_zero(), fr_zero(), fr_zero()};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
    fr_t s0 = fr_zero(), s2 = fr_zero(), s3 = fr_zero();
    for (int k = blockIdx.y; k < ncirc; k += gridDim.y) {
      const fr_t* a = A[k];
      const fr_t* b = B[k];
      fr_t a0 = ld_fr_stream(a + i), a1 = ld_fr_stream(a + half + i);
      if (scale) {
        a0 = fr_mul(cf.v[k], a0);
        a1 = fr_mul(cf.v[k], a1);
      }
      const fr_t b0 = ld_fr_stream(b + i), b1 = ld_fr_stream(b + half + i);
      const fr_t da = fr_sub(a1, a0), db = fr_sub(b1, b0);
      const fr_t a2 = fr_add(a1, da), b2 = fr_add(b1, db);
      s0 = fr_add(s0, fr_mul(a0, b0));
      s2 = fr_add(s2, fr_mul(a2, b2));
      s3 = fr_add(s3, fr_mul(fr_add(a2, da), fr_add(b2, db)));
    }
    const fr_t c0 = ld_fr(Ceq + i), c1 = ld_fr(Ceq + half + i), dc = fr_sub(c1, c0), c2 = fr_add(c1, dc);
    acc[0] = fr_add(acc[0], fr_mul(s0, c0));
    acc[1] = fr_add(acc[1], fr_mul(s2, c2));
    acc[2] = fr_add(acc[2], fr_mul(s3, fr_add(c2, dc)));
  }
  block_sum_fr<3>(acc, scratch);
  const int total = gridDim.x * gridDim.y;
  finalize_block<3>(fin, acc, 0, blockIdx.y * gridDim.x + blockIdx.x, total, 3, total);
}
// h = number of bound outputs per polynomial (current length / 2), must be >= 2.
__global__ void __launch_bounds__(kThreads)
    sc_bind_eval_cubic_comb_kernel(fr_t* const* A, fr_t* const* B, const fr_t* Cin, fr_t* Cout, size_t h, fr_t r, int ncirc,
                                   CubicCoeffs cf, int scale, Finalize fin) {
  __shared__ fr_t scratch[3 * kThreads / 32];
  const size_t q = h / 2;
  fr_t acc[3] = {fr_zero(), fr_zero(), fr_zero()};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < q; i += (size_t)gridDim.x * blockDim.x) {
    fr_t s0 = fr_zero(), s2 = fr_zero(), s3 = fr_zero();
    fr_t lo, hi;
    for (int k = blockIdx.y; k < ncirc; k += gridDim.y) {
      fr_t* a = A[k];
      fr_t* b = B[k];
      lo = ld_fr_stream(a + i); hi = ld_fr_stream(a + i + h);
      fr_t a0 = fr_add(lo, fr_mul(r, fr_sub(hi, lo)));
      lo = ld_fr_stream(a + i + q); hi = ld_fr_stream(a + i + q + h);
      fr_t a1 = fr_add(lo, fr_mul(r, fr_sub(hi, lo)));
      if (scale) {
        a0 = fr_mul(cf.v[k], a0);
        a1 = fr_mul(cf.v[k], a1);
      }
      st_fr(a + i, a0);
      st_fr(a + i + q, a1);
      lo = ld_fr_stream(b + i); hi = ld_fr_stream(b + i + h);
      const fr_t b0 = fr_add(lo, fr_mul(r, fr_sub(hi, lo)));
      lo = ld_fr_stream(b + i + q); hi = ld_fr_stream(b + i + q + h);
      const fr_t b1 = fr_add(lo, fr_mul(r, fr_sub(hi, lo)));
      st_fr(b + i, b0);
      st_fr(b + i + q, b1);
      const fr_t da = fr_sub(a1, a0), db = fr_sub(b1, b0);
      const fr_t a2 = fr_add(a1, da), b2 = fr_add(b1, db);
      s0 = fr_add(s0, fr_mul(a0, b0));
      s2 = fr_add(s2, fr_mul(a2, b2));
      s3 = fr_add(s3, fr_mul(fr_add(a2, da), fr_add(b2, db)));
    }
    lo = ld_fr(Cin + i); hi = ld_fr(Cin + i + h);
    const fr_t c0 = fr_add(lo, fr_mul(r, fr_sub(hi, lo)));
    lo = ld_fr(Cin + i + q); hi = ld_fr(Cin + i + q + h);
    const fr_t c1 = fr_add(lo, fr_mul(r, fr_sub(hi, lo)));
    if (blockIdx.y == 0) {
      st_fr(Cout + i, c0);
      st_fr(Cout + i + q, c1);
    }
    const fr_t dc = fr_sub(c1, c0), c2 = fr_add(c1, dc);
    acc[0] = fr_add(acc[0], fr_mul(s0, c0));
    acc[1] = fr_add(acc[1], fr_mul(s2, c2));
    acc[2] = fr_add(acc[2], fr_mul(s3, fr_add(c2, dc)));
  }
  block_sum_fr<3>(acc, scratch);
  const int total = gridDim.x * gridDim.y;
  finalize_block<3>(fin, acc, 0, blockIdx.y * gridDim.x + blockIdx.x, total, 3, total);
}
// Latency-oriented variant of the two kernels above for the small and medium rounds (most of the ~300 rounds of
// a grand-product argument move a few KB: what the host waits for is the dependent chain inside one thread,
// 12 field multiplications ~ 4.5 us on a lone warp).  FOUR lanes per (circuit k, pair index i):
//   lane 0 / 1 / 2 binds its side (A_k / B_k / eq) at i and i+q (2 multiplications, do_bind != 0) and forms
//   the side's values at t = 0, 2, 3; three quad shuffles hand lane t the three factors of its evaluation
//   point, which it multiplies (2 multiplications): 4 dependent multiplications instead of 12.
// Sums over i: xor-shuffles inside the warp, shared memory across warps, then either a direct tagged
// publication (single CTA: no ticket, no fence) or the usual Finalize last-CTA stage.
//   do_bind = 1: arrays hold 4q elements, pairs (i, i+2q) are bound with r into (i), then evaluated as (i, i+q)
//   do_bind = 0: arrays hold 2q elements, evaluated as (i, i+q)               (q a power of two)
__global__ void __launch_bounds__(1024)
    sc_cubic_quad_kernel(fr_t* const* A, fr_t* const* B, const fr_t* Cin, fr_t* Cout, size_t q, int lg_q, int do_bind,
                         fr_t r, int ncirc, CubicCoeffs cf, int scale, int comb, Finalize fin) {
  // cf / scale / comb: see the combined kernels above — scale: lane 0 multiplies its side by coeff_k (stored when
  // binding); comb: the CTA yields 3 values (summed over all its circuits) instead of 3 per circuit
  __shared__ fr_t s_part[256 * 3];
  __shared__ int s_last;
  const int tid = threadIdx.x, role = tid & 3, lane = tid & 31;
  const int upb = blockDim.x >> 2;  // (circuit, pair) units per CTA
  const size_t U = (size_t)blockIdx.x * upb + (tid >> 2), total = (size_t)ncirc << lg_q;
  const bool valid = U < total;
  const int k = valid ? (int)(U >> lg_q) : 0;
  const size_t i = U & (q - 1), h = 2 * q;
  fr_t x0 = fr_zero(), x1 = fr_zero();
  if (valid && role < 3) {
    fr_t* src = role == 0 ? A[k] : role == 1 ? B[k] : const_cast<fr_t*>(Cin);
    if (do_bind) {
      fr_t lo = ld_fr(src + i), hi = ld_fr(src + i + h);
      x0 = fr_add(lo, fr_mul(r, fr_sub(hi, lo)));
      lo = ld_fr(src + i + q);
      hi = ld_fr(src + i + q + h);
      x1 = fr_add(lo, fr_mul(r, fr_sub(hi, lo)));
    } else {
      x0 = ld_fr(src + i);
      x1 = ld_fr(src + i + q);
    }
    if (scale && role == 0) {
      x0 = fr_mul(cf.v[k], x0);
      x1 = fr_mul(cf.v[k], x1);
    }
    if (do_bind) {
      fr_t* dst = role == 2 ? (k == 0 ? Cout : nullptr) : src;
      if (dst) {
        st_fr(dst + i, x0);
        st_fr(dst + i + q, x1);
      }
    }
  }
  // this side at t = 0, 2, 3
  fr_t e0 = x0, d = fr_sub(x1, x0), e2 = fr_add(x1, d), e3 = fr_add(e2, d);
  // round j: side s offers its value at t = (s + j) % 3; lane t reads side (t - j) mod 3 -> that side at t
  fr_t P;
  const int qbase = lane & ~3;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const int sel = (role + j) % 3;
    const fr_t offer = sel == 0 ? e0 : (sel == 1 ? e2 : e3);
    const int src_lane = qbase + (role + 3 - j) % 3;
    fr_t got;
#pragma unroll
    for (int l = 0; l < 8; l++) got.v[l] = __shfl_sync(0xffffffffu, offer.v[l], src_lane);
    P = j == 0 ? got : fr_mul(P, got);
  }
  if (!valid || role == 3) P = fr_zero();
  // sum over the pair indices of a circuit: gq = min(q, 8) consecutive units of a warp belong to one circuit
  // (combined: all 8 units of the warp, whatever their circuit)
  const int gq = comb ? 8 : (q < 8 ? (int)q : 8);
  for (int off = 1; off < gq; off <<= 1) {
    fr_t o;
#pragma unroll
    for (int l = 0; l < 8; l++) o.v[l] = __shfl_xor_sync(0xffffffffu, P.v[l], off * 4);
    P = fr_add(P, o);
  }
  const int unit = tid >> 2;
  if ((unit & (gq - 1)) == 0 && role < 3) s_part[(unit / gq) * 3 + role] = P;
  __syncthreads();
  // block outputs: cpb circuits x 3 values, each the sum of gpc group partials (combined: 3 values, all warps)
  const int cpb = q >= (size_t)upb ? 1 : upb >> lg_q;
  const int gpc = (int)((q >= (size_t)upb ? (size_t)upb : q) / gq);
  const int bpv = comb ? (int)gridDim.x : (q >= (size_t)upb ? (int)(q / upb) : 1);  // CTAs per value
  int v = -1;
  fr_t val = fr_zero();
  if (comb) {
    if (tid < 3) {
      const int nw = blockDim.x >> 5;
      for (int w = 0; w < nw; w++) val = fr_add(val, s_part[w * 3 + tid]);
      v = tid;
    }
  } else if (tid < cpb * 3) {
    const int cl = tid / 3, t = tid - 3 * cl;
    const int kk = q >= (size_t)upb ? (int)(blockIdx.x / bpv) : (int)blockIdx.x * cpb + cl;
    if (kk < ncirc) {
      for (int w = 0; w < gpc; w++) val = fr_add(val, s_part[(cl * gpc + w) * 3 + t]);
      v = kk * 3 + t;
    }
  }
  if (gridDim.x == 1) {
    if (v >= 0) finalize_publish(fin, v, val);
    return;
  }
  if (v >= 0) {
    fin.partial[(size_t)v * bpv + (blockIdx.x % bpv)] = val;
    __threadfence();
  }
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(fin.counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  finalize_last_stage(fin, bpv, comb ? 3 : 3 * ncirc);
}
static constexpr size_t kQuadMaxQ = 2048;  // beyond this the rounds are throughput-bound: thread-per-pair kernels
static void launch_cubic_quad(fr_t* const* d_A, fr_t* const* d_B, const fr_t* Cin, fr_t* Cout, int ncirc, size_t q,
                              int do_bind, const fr_t& r, const CubicCoeffs& cf, int scale, int comb, const Finalize& fin,
                              cudaStream_t st) {
  int lg_q = 0;
  while (((size_t)1 << lg_q) < q) lg_q++;
  const size_t threads = 4 * (size_t)ncirc * q;
  if (threads <= 1024) {
    unsigned t = (unsigned)((threads + 31) / 32 * 32);
    sc_cubic_quad_kernel<<<1, t, 0, st>>>(d_A, d_B, Cin, Cout, q, lg_q, do_bind, r, ncirc, cf, scale, comb, fin);
    LB_LAUNCH_CHECK();
  } else {
    unsigned blocks = (unsigned)(((size_t)ncirc * q + 63) / 64);
    sc_cubic_quad_kernel<<<blocks, 256, 0, st>>>(d_A, d_B, Cin, Cout, q, lg_q, do_bind, r, ncirc, cf, scale, comb, fin);
    LB_LAUNCH_CHECK();
  }
}
// circuits are strided over blockIdx.y: enough CTAs for ~2 per SM even when a round has few pairs
static dim3 comb_grid(size_t pairs, int ncirc) {
  int bx = grid_for(pairs, kThreads, kMaxBlocks);
  int gy = (2 * kNumSMs + bx - 1) / bx;
  if (gy > ncirc) gy = ncirc;
  if (gy < 1) gy = 1;
  return dim3(bx, gy);
}
void launch_sumcheck_bind_eval_cubic_comb(fr_t* const* d_A, fr_t* const* d_B, const fr_t* Cin, fr_t* Cout, int ncirc, size_t h,
                                          const fr_t& r, const CubicCoeffs& cf, int scale, const Finalize& fin, cudaStream_t st) {
  size_t q = h / 2;
  if (q <= kQuadMaxQ && (q & (q - 1)) == 0) return launch_cubic_quad(d_A, d_B, Cin, Cout, ncirc, q, 1, r, cf, scale, 1, fin, st);
  sc_bind_eval_cubic_comb_kernel<<<comb_grid(q, ncirc), kThreads, 0, st>>>(d_A, d_B, Cin, Cout, h, r, ncirc, cf, scale, fin);
  LB_LAUNCH_CHECK();
}
void launch_sumcheck_eval_cubic_comb(fr_t* const* d_A, fr_t* const* d_B, const fr_t* Ceq, int ncirc, size_t half,
                                     const CubicCoeffs& cf, int scale, const Finalize& fin, cudaStream_t st) {
  if (half <= kQuadMaxQ && (half & (half - 1)) == 0)
    return launch_cubic_quad(d_A, d_B, Ceq, nullptr, ncirc, half, 0, fr_zero(), cf, scale, 1, fin, st);
  sc_eval_cubic_comb_kernel<<<comb_grid(half, ncirc), kThreads, 0, st>>>(d_A, d_B, Ceq, half, ncirc, cf, scale, fin);
  LB_LAUNCH_CHECK();
}
// per-circuit outputs (e0, e2, e3)_k, no batching coefficients: the per-loop C-ABI entry lasso_sumcheck_round_cubic
void launch_sumcheck_eval_cubic(fr_t* const* d_A, fr_t* const* d_B, const fr_t* Ceq, int ncirc, size_t half,
                                const Finalize& fin, cudaStream_t st) {
  if (half <= kQuadMaxQ && (half & (half - 1)) == 0) {
    CubicCoeffs none;
    return launch_cubic_quad(d_A, d_B, Ceq, nullptr, ncirc, half, 0, fr_zero(), none, 0, 0, fin, st);
  }
  int per = kMaxBlocks / ncirc;
  if (per < 1) per = 1;
  int bx = grid_for(half, kThreads, per);
  dim3 grid(bx, ncirc);
  sc_eval_cubic_kernel<<<grid, kThreads, 0, st>>>(d_A, d_B, Ceq, half, fin);
  LB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------ K5
__device__ __forceinline__ uint32_t subtable_value(int kind, int sub, uint32_t idx, int log_m, int log_r) {
  if (kind == STRAT_RANGE) {  // range_check.rs:15-34
    if (sub == 0) return idx;
    if (sub == 1) return idx < (1u << (log_r % log_m)) ? idx : 0u;
    return 0u;
  }
  int bits = log_m / 2;  // utils/mod.rs:82-89 split_bits: (high, low)
  uint32_t lhs = (idx >> bits) & ((1u << bits) - 1), rhs = idx & ((1u << bits) - 1);
  switch (kind) {
    case STRAT_AND: return lhs & rhs;
    case STRAT_OR: return lhs | rhs;
    case STRAT_XOR: return lhs ^ rhs;
    default: return sub == 0 ? (lhs < rhs) : (lhs == rhs);  // lt.rs:16-30
  }
}
__global__ void __launch_bounds__(kThreads)
    materialize_kernel(int kind, int nsub, int log_m, int log_r, fr_t* tables_fr, uint32_t* tables_u32) {
  size_t M = (size_t)1 << log_m, n = M * nsub;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t v = subtable_value(kind, (int)(i >> log_m), (uint32_t)(i & (M - 1)), log_m, log_r);
    if (tables_u32) tables_u32[i] = v;
    if (tables_fr) st_fr(tables_fr + i, fr_from_u64(v));
  }
}
void launch_materialize_subtables(const Strategy& S, fr_t* tables_fr, uint32_t* tables_u32, cudaStream_t st) {
  size_t n = (size_t)S.M() * S.num_subtables();
  materialize_kernel<<<grid_for(n), kThreads, 0, st>>>(S.kind, S.num_subtables(), S.log_m, S.log_r, tables_fr,
                                                       tables_u32);
  LB_LAUNCH_CHECK();
}
struct GatherMap {
  int sub[32], dim[32];
};
// subtables/mod.rs:78-92: E_k[j] = T_sub(k)[nz_dim(k)[j]]; 32 B written per (memory, lookup)
__global__ void __launch_bounds__(kThreads)
    gather_kernel(GatherMap map, int log_m, const fr_t* tables_fr, const uint32_t* tables_u32, const uint32_t* nz,
                  size_t s, fr_t* E_fr, size_t E_stride, uint32_t* E_u32) {
  int k = blockIdx.y;
  const uint32_t* idx = nz + (size_t)map.dim[k] * s;
  size_t toff = (size_t)map.sub[k] << log_m;
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < s; j += (size_t)gridDim.x * blockDim.x) {
    uint32_t a = idx[j];
    if (E_fr) st_fr(E_fr + (size_t)k * E_stride + j, ld_fr(tables_fr + toff + a));
    if (E_u32) E_u32[(size_t)k * s + j] = tables_u32[toff + a];
  }
}
void launch_gather_lookup_polys(const Strategy& S, const fr_t* tables_fr, const uint32_t* tables_u32,
                                const uint32_t* nz, size_t s, fr_t* E_fr, size_t E_stride, uint32_t* E_u32,
                                cudaStream_t st) {
  GatherMap map;
  for (int k = 0; k < S.num_memories(); k++) {
    map.sub[k] = S.memory_to_subtable_index(k);
    map.dim[k] = S.memory_to_dimension_index(k);
  }
  dim3 grid(grid_for(s, kThreads, kMaxBlocks / S.num_memories() + 1), S.num_memories());
  gather_kernel<<<grid, kThreads, 0, st>>>(map, S.log_m, tables_fr, tables_u32, nz, s, E_fr, E_stride, E_u32);
  LB_LAUNCH_CHECK();
}
__global__ void __launch_bounds__(kThreads) from_u32_kernel(const uint32_t* in, fr_t* out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    st_fr(out + i, fr_from_u64(in[i]));
}
void launch_from_u32(const uint32_t* in, fr_t* out, size_t n, cudaStream_t st) {
  if (n) from_u32_kernel<<<grid_for(n), kThreads, 0, st>>>(in, out, n);
}
void launch_fill_zero(fr_t* out, size_t n, cudaStream_t st) {
  if (n) cudaMemsetAsync(out, 0, n * sizeof(fr_t), st);
}

// ------------------------------------------------------------------------------------ K7
// dense_mlpoly.rs:228-235 + utils/mod.rs:63-73 with the eq table shared by all polynomials
__global__ void __launch_bounds__(kThreads)
    multi_dot_kernel(const fr_t* base, size_t stride, const fr_t* eq, size_t n, fr_t* partial) {
  __shared__ fr_t scratch[kThreads / 32];
  const fr_t* P = base + (size_t)blockIdx.y * stride;
  fr_t acc[1] = {fr_zero()};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    acc[0] = fr_add(acc[0], fr_mul(ld_fr_stream(P + i), ld_fr(eq + i)));
  block_sum_fr<1>(acc, scratch);
  if (threadIdx.x == 0) partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = acc[0];
}
void launch_multi_dot(const fr_t* base, size_t stride, int npolys, const fr_t* eq, size_t n, fr_t* partial,
                      fr_t* out, cudaStream_t st) {
  int per = kMaxBlocks / npolys;
  if (per < 1) per = 1;
  int bx = grid_for(n, kThreads, per);
  dim3 grid(bx, npolys);
  multi_dot_kernel<<<grid, kThreads, 0, st>>>(base, stride, eq, n, partial);
  LB_LAUNCH_CHECK();
  reduce_partials_kernel<<<npolys, kThreads, 0, st>>>(partial, bx, out);
  LB_LAUNCH_CHECK();
}

// dense_mlpoly.rs:183-207: LZ[i] = sum_j L[j] Z[j*R + i].  Thread = column (coalesced across the
// warp), rows split into chunks over blockIdx.y, second pass sums the chunk partials.
static constexpr int kBoundChunks = 64;
int bound_max_chunks() { return kBoundChunks; }
__global__ void __launch_bounds__(kThreads)
    bound_kernel(const fr_t* Z, const fr_t* L, size_t L_size, size_t R_size, size_t rows_per_chunk, fr_t* partial) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R_size) return;
  size_t j0 = (size_t)blockIdx.y * rows_per_chunk, j1 = j0 + rows_per_chunk;
  if (j1 > L_size) j1 = L_size;
  fr_t acc = fr_zero();
  for (size_t j = j0; j < j1; j++) acc = fr_add(acc, fr_mul(L[j], ld_fr_stream(Z + j * R_size + i)));
  st_fr(partial + (size_t)blockIdx.y * R_size + i, acc);
}
__global__ void __launch_bounds__(kThreads) bound_reduce_kernel(const fr_t* partial, int chunks, size_t R_size, fr_t* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R_size) return;
  fr_t acc = fr_zero();
  for (int c = 0; c < chunks; c++) acc = fr_add(acc, ld_fr(partial + (size_t)c * R_size + i));
  st_fr(out + i, acc);
}
void launch_bound(const fr_t* Z, const fr_t* L, size_t L_size, size_t R_size, fr_t* partial, fr_t* out,
                  cudaStream_t st) {
  int chunks = (int)(L_size < (size_t)kBoundChunks ? L_size : (size_t)kBoundChunks);
  size_t rows_per_chunk = (L_size + chunks - 1) / chunks;
  dim3 grid((unsigned)((R_size + kThreads - 1) / kThreads), chunks);
  bound_kernel<<<grid, kThreads, 0, st>>>(Z, L, L_size, R_size, rows_per_chunk, partial);
  LB_LAUNCH_CHECK();
  bound_reduce_kernel<<<(unsigned)((R_size + kThreads - 1) / kThreads), kThreads, 0, st>>>(partial, chunks, R_size, out);
  LB_LAUNCH_CHECK();
}

// ---- the same two reductions for INTEGER-valued polynomials (dim, read, final, E: everything the prover commits to and
// opens is an index, a counter or a table value < 2^32, kept as a u32 mirror next to the field form).  A field element
// times a 32-bit integer is 8 IMAD.WIDE instead of a ~235-instruction Montgomery product, and the sum can be carried
// as a plain 320-bit integer (X = sum_j L_j * z_j < 2^288 * #terms) and reduced ONCE: L_j is stored as L_j*R mod l, so
// X mod l is already the Montgomery form of the result.  Reads 4 B per element instead of 32 B.
struct wide_t {
  uint32_t v[10];
};
__device__ __forceinline__ void wide_zero(wide_t& a) {
#pragma unroll
  for (int l = 0; l < 10; l++) a.v[l] = 0;
}
__device__ __forceinline__ void wide_mad(wide_t& acc, const fr_t& a, uint32_t z) {  // acc += a * z
  uint64_t carry = 0;
#pragma unroll
  for (int l = 0; l < 8; l++) {
    const uint64_t t = (uint64_t)a.v[l] * z + acc.v[l] + carry;
    acc.v[l] = (uint32_t)t;
    carry = t >> 32;
  }
  const uint64_t t = (uint64_t)acc.v[8] + carry;
  acc.v[8] = (uint32_t)t;
  acc.v[9] += (uint32_t)(t >> 32);
}
// X mod l for X < 2^320, as a field element: X = X_lo + 2^256 * X_hi;  X_lo mod l through two Montgomery products
// (x -> x*R -> x), X_hi * 2^256 mod l = the Montgomery form of the 64-bit integer X_hi
__device__ __forceinline__ fr_t wide_reduce(const wide_t& a) {
  fr_t lo;
#pragma unroll
  for (int l = 0; l < 8; l++) lo.v[l] = a.v[l];
  const fr_t lo_mod = fr_to_canonical(fr_from_raw_int(lo));
  return fr_add(lo_mod, fr_from_u64((uint64_t)a.v[8] | ((uint64_t)a.v[9] << 32)));
}
// LZ[i] = sum_j L[j] z[j*R + i]: thread = column, rows split into chunks over blockIdx.y (<= 2^20 rows per chunk)
__global__ void __launch_bounds__(kThreads)
    bound_u32_kernel(const uint32_t* Z, const fr_t* L, size_t L_size, size_t R_size, size_t rows_per_chunk, fr_t* partial) {
  __shared__ fr_t sL[64];
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t j0 = (size_t)blockIdx.y * rows_per_chunk;
  size_t j1 = j0 + rows_per_chunk;
  if (j1 > L_size) j1 = L_size;
  wide_t acc;
  wide_zero(acc);
  for (size_t jb = j0; jb < j1; jb += 64) {  // the row weights of 64 rows at a time through shared memory
    __syncthreads();
    if (threadIdx.x < 64 && jb + threadIdx.x < j1) sL[threadIdx.x] = ld_fr(L + jb + threadIdx.x);
    __syncthreads();
    const size_t je = jb + 64 < j1 ? jb + 64 : j1;
    if (i < R_size)
      for (size_t j = jb; j < je; j++) wide_mad(acc, sL[j - jb], Z[j * R_size + i]);
  }
  if (i < R_size) st_fr(partial + (size_t)blockIdx.y * R_size + i, wide_reduce(acc));
}
void launch_bound_u32(const uint32_t* Z, const fr_t* L, size_t L_size, size_t R_size, fr_t* partial, fr_t* out, cudaStream_t st) {
  int chunks = (int)(L_size < (size_t)kBoundChunks ? L_size : (size_t)kBoundChunks);
  size_t rows_per_chunk = (L_size + chunks - 1) / chunks;
  if (rows_per_chunk > ((size_t)1 << 20)) throw std::runtime_error("bound_u32: too many rows per chunk");
  dim3 grid((unsigned)((R_size + kThreads - 1) / kThreads), chunks);
  bound_u32_kernel<<<grid, kThreads, 0, st>>>(Z, L, L_size, R_size, rows_per_chunk, partial);
  LB_LAUNCH_CHECK();
  bound_reduce_kernel<<<(unsigned)((R_size + kThreads - 1) / kThreads), kThreads, 0, st>>>(partial, chunks, R_size, out);
  LB_LAUNCH_CHECK();
}
// out[k] = <z_k, eq>, z_k = base + k*stride (u32), k < npolys
__global__ void __launch_bounds__(kThreads)
    multi_dot_u32_kernel(const uint32_t* base, size_t stride, const fr_t* eq, size_t n, fr_t* partial) {
  __shared__ fr_t scratch[kThreads / 32];
  const uint32_t* P = base + (size_t)blockIdx.y * stride;
  wide_t w;
  wide_zero(w);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    wide_mad(w, ld_fr(eq + i), P[i]);  // a thread adds at most n / (gridDim.x * 256) < 2^32 terms
  fr_t acc[1] = {wide_reduce(w)};
  block_sum_fr<1>(acc, scratch);
  if (threadIdx.x == 0) partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = acc[0];
}
void launch_multi_dot_u32(const uint32_t* base, size_t stride, int npolys, const fr_t* eq, size_t n, fr_t* partial, fr_t* out,
                          cudaStream_t st) {
  int per = kMaxBlocks / npolys;
  if (per < 1) per = 1;
  int bx = grid_for(n, kThreads, per);
  dim3 grid(bx, npolys);
  multi_dot_u32_kernel<<<grid, kThreads, 0, st>>>(base, stride, eq, n, partial);
  LB_LAUNCH_CHECK();
  reduce_partials_kernel<<<npolys, kThreads, 0, st>>>(partial, bx, out);
  LB_LAUNCH_CHECK();
}

// memory_checking.rs:249-252: hash(a, v, t) = t*gamma^2 + v*gamma + a - tau
__global__ void __launch_bounds__(kThreads)
    fp_mem_kernel(const fr_t* table, const fr_t* final_fr, size_t M, int G, int g, fr_t gamma, fr_t gamma2, fr_t tau,
                  fr_t* out_init, fr_t* out_final) {
  // M = local cells; local cell i is global address i*G + g (low-bit partition); `table` is the full table
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (size_t)gridDim.x * blockDim.x) {
    size_t gi = i * G + g;
    fr_t h0 = fr_sub(fr_add(fr_mul(ld_fr(table + gi), gamma), fr_from_u64(gi)), tau);  // ts = 0
    st_fr(out_init + i, h0);
    st_fr(out_final + i, fr_add(h0, fr_mul(ld_fr(final_fr + i), gamma2)));
  }
}
__global__ void __launch_bounds__(kThreads)
    fp_ops_kernel(const fr_t* dim_fr, const fr_t* E_fr, const fr_t* read_fr, size_t s, fr_t gamma, fr_t gamma2,
                  fr_t tau, fr_t* out_read, fr_t* out_write) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < s; i += (size_t)gridDim.x * blockDim.x) {
    fr_t av = fr_sub(fr_add(fr_mul(ld_fr_stream(E_fr + i), gamma), ld_fr_stream(dim_fr + i)), tau);
    fr_t hr = fr_add(av, fr_mul(ld_fr_stream(read_fr + i), gamma2));
    st_fr(out_read + i, hr);
    st_fr(out_write + i, fr_add(hr, gamma2));  // write ts = read ts + 1
  }
}
void launch_gp_fingerprints_mem(const fr_t* table, const fr_t* final_fr, size_t M_local, int G, int g,
                                const fr_t& gamma, const fr_t& tau, fr_t* out_init, fr_t* out_final, cudaStream_t st) {
  fp_mem_kernel<<<grid_for(M_local), kThreads, 0, st>>>(table, final_fr, M_local, G, g, gamma, fr_sqr(gamma), tau,
                                                        out_init, out_final);
  LB_LAUNCH_CHECK();
}
void launch_gp_fingerprints_ops(const fr_t* dim_fr, const fr_t* E_fr, const fr_t* read_fr, size_t s,
                                const fr_t& gamma, const fr_t& tau, fr_t* out_read, fr_t* out_write,
                                cudaStream_t st) {
  fp_ops_kernel<<<grid_for(s), kThreads, 0, st>>>(dim_fr, E_fr, read_fr, s, gamma, fr_sqr(gamma), tau, out_read,
                                                  out_write);
  LB_LAUNCH_CHECK();
}
// grand_product.rs:20-36 with the layer stored contiguously as [left | right]
__global__ void __launch_bounds__(kThreads) product_layer_kernel(const fr_t* in, fr_t* out, size_t n_out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (size_t)gridDim.x * blockDim.x)
    st_fr(out + i, fr_mul(ld_fr_stream(in + i), ld_fr_stream(in + n_out + i)));
}
void launch_product_layer(const fr_t* in, fr_t* out, size_t n_out, cudaStream_t st) {
  product_layer_kernel<<<grid_for(n_out), kThreads, 0, st>>>(in, out, n_out);
  LB_LAUNCH_CHECK();
}

// All product trees of one size at once (single GPU).  A tree is one contiguous array: layer 0 (N elements),
// then layer 1 (N/2), ...; layer k+1[i] = layer k[i] * layer k[i + len/2].  One launch per layer for every
// tree (blockIdx.y) while the layers are large, then ONE CTA per tree walks the remaining small layers with
// a barrier in between and publishes the two elements of the top layer (grand_product.rs:60-65 `evaluate`)
// as tagged values 2*slot0 + 2*tree + {0, 1}: ~16 launches per proof instead of ~270 + 16 small copies.
// stop_len = 2: a whole tree (single GPU).  stop_len = 1: the tree of a low-bit SHARD (N = local length) — the walk ends
// with the rank's single element of the layer of global length G, published as value slot0 + tree.
__global__ void __launch_bounds__(kThreads) product_layers_kernel(TreePtrs trees, size_t in_off, size_t n_out) {
  const fr_t* in = trees.p[blockIdx.y] + in_off;
  fr_t* out = trees.p[blockIdx.y] + in_off + 2 * n_out;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (size_t)gridDim.x * blockDim.x)
    st_fr(out + i, fr_mul(ld_fr(in + i), ld_fr(in + n_out + i)));
}
__global__ void __launch_bounds__(1024)
    product_tail_kernel(TreePtrs trees, size_t off, size_t len, int slot0, int stop_len, Finalize fin) {
  fr_t* base = trees.p[blockIdx.x];
  while (len > (size_t)stop_len) {
    const size_t n_out = len / 2;
    for (size_t i = threadIdx.x; i < n_out; i += blockDim.x)
      st_fr(base + off + len + i, fr_mul(ld_fr(base + off + i), ld_fr(base + off + n_out + i)));
    __syncthreads();
    off += len;
    len = n_out;
  }
  if ((int)threadIdx.x < stop_len)
    finalize_publish(fin, stop_len * (slot0 + (int)blockIdx.x) + (int)threadIdx.x, ld_fr(base + off + threadIdx.x));
}
void launch_product_trees(const TreePtrs& trees, int ntrees, size_t N, int slot0, int stop_len, const Finalize& fin,
                          cudaStream_t st) {
  size_t off = 0, len = N;
  while (len > 4096) {
    const size_t n_out = len / 2;
    dim3 grid(grid_for(n_out, kThreads, kMaxBlocks / ntrees + 1), ntrees);
    product_layers_kernel<<<grid, kThreads, 0, st>>>(trees, off, n_out);
    LB_LAUNCH_CHECK();
    off += len;
    len = n_out;
  }
  product_tail_kernel<<<ntrees, 1024, 0, st>>>(trees, off, len, slot0, stop_len, fin);
  LB_LAUNCH_CHECK();
}
int product_trees_launches(size_t N) {
  int n = 1;
  for (size_t len = N; len > 4096; len /= 2) n++;
  return n;
}
// last round of a batched cubic sumcheck (one element pair left per array): bind the 2*ncirc heads with r in
// place and publish them — they are the layer's claims (grand_product.rs:139-150)
__global__ void bind_heads_kernel(fr_t* const* AB, int n, fr_t r, Finalize fin) {
  const int k = threadIdx.x;
  if (k >= n) return;
  fr_t* x = AB[k];
  const fr_t lo = ld_fr(x), hi = ld_fr(x + 1);
  const fr_t v = fr_add(lo, fr_mul(r, fr_sub(hi, lo)));
  st_fr(x, v);
  finalize_publish(fin, k, v);
}
void launch_bind_heads(fr_t* const* d_AB, int n, const fr_t& r, const Finalize& fin, cudaStream_t st) {
  bind_heads_kernel<<<1, (n + 31) / 32 * 32, 0, st>>>(d_AB, n, r, fin);
  LB_LAUNCH_CHECK();
}

// ---- Bulletproofs scalar-side helpers (bullet.rs:73-134) ----
__global__ void __launch_bounds__(kThreads) fold_ab_kernel(fr_t* a, fr_t* b, size_t h, fr_t u, fr_t uinv) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < h; i += (size_t)gridDim.x * blockDim.x) {
    fr_t aL = ld_fr(a + i), aR = ld_fr(a + h + i), bL = ld_fr(b + i), bR = ld_fr(b + h + i);
    st_fr(a + i, fr_add(fr_mul(aL, u), fr_mul(uinv, aR)));
    st_fr(b + i, fr_add(fr_mul(bL, uinv), fr_mul(u, bR)));
  }
}
void launch_fold_ab(fr_t* a, fr_t* b, size_t h, const fr_t& u, const fr_t& uinv, cudaStream_t st) {
  fold_ab_kernel<<<grid_for(h), kThreads, 0, st>>>(a, b, h, u, uinv);
  LB_LAUNCH_CHECK();
}
__global__ void __launch_bounds__(kThreads) cross_ip_kernel(const fr_t* a, const fr_t* b, size_t h, fr_t* partial) {
  __shared__ fr_t scratch[2 * kThreads / 32];
  fr_t acc[2] = {fr_zero(), fr_zero()};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < h; i += (size_t)gridDim.x * blockDim.x) {
    acc[0] = fr_add(acc[0], fr_mul(ld_fr(a + i), ld_fr(b + h + i)));
    acc[1] = fr_add(acc[1], fr_mul(ld_fr(a + h + i), ld_fr(b + i)));
  }
  block_sum_fr<2>(acc, scratch);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = acc[0];
    partial[gridDim.x + blockIdx.x] = acc[1];
  }
}
void launch_cross_inner_products(const fr_t* a, const fr_t* b, size_t h, fr_t* partial, fr_t* out, cudaStream_t st) {
  int bx = grid_for(h, kThreads, 64);
  cross_ip_kernel<<<bx, kThreads, 0, st>>>(a, b, h, partial);
  LB_LAUNCH_CHECK();
  reduce_partials_kernel<<<2, kThreads, 0, st>>>(partial, bx, out);
  LB_LAUNCH_CHECK();
}
__global__ void __launch_bounds__(kThreads)
    expand_weights_kernel(const fr_t* w, fr_t* w_out, size_t n_in, fr_t u, fr_t uinv) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_in; t += (size_t)gridDim.x * blockDim.x) {
    fr_t x = ld_fr(w + t);
    st_fr(w_out + 2 * t, fr_mul(x, uinv));
    st_fr(w_out + 2 * t + 1, fr_mul(x, u));
  }
}
void launch_expand_weights(const fr_t* w, fr_t* w_out, size_t n_in, const fr_t& u, const fr_t& uinv, cudaStream_t st) {
  expand_weights_kernel<<<grid_for(n_in), kThreads, 0, st>>>(w, w_out, n_in, u, uinv);
  LB_LAUNCH_CHECK();
}
// Round with current vector length m (half h = m/2) over n original generators, weights w[t], t < n/m.
// Global column j = t*m + pos:   sL[j] = a[pos-h] * w[t] for pos >= h (else 0),
//                                sR[j] = a[h+pos] * w[t] for pos <  h (else 0).
// Sharded over G GPUs this rank owns the columns j = j'*G + g (n_loc of them).  `a` is either this rank's
// low-bit shard of the folded vector (a_rep = 0, valid while m >= 2G: element p lives at p / G) or the
// replicated full vector of the tail rounds (a_rep = 1).
__global__ void __launch_bounds__(kThreads)
    bullet_scalars_kernel(const fr_t* a, const fr_t* w, size_t n_loc, size_t m, int G, int g, int a_rep, fr_t* sL,
                          fr_t* sR) {
  size_t h = m / 2;
  for (size_t jl = (size_t)blockIdx.x * blockDim.x + threadIdx.x; jl < n_loc; jl += (size_t)gridDim.x * blockDim.x) {
    size_t j = jl * G + g;
    size_t t = j / m, pos = j % m;
    fr_t wt = ld_fr(w + t);
    if (pos >= h) {
      size_t idx = pos - h;
      st_fr(sL + jl, fr_mul(ld_fr(a + (a_rep ? idx : idx / G)), wt));
      st_fr(sR + jl, fr_zero());
    } else {
      size_t idx = h + pos;
      st_fr(sL + jl, fr_zero());
      st_fr(sR + jl, fr_mul(ld_fr(a + (a_rep ? idx : idx / G)), wt));
    }
  }
}
void launch_bullet_scalars(const fr_t* a, const fr_t* w, size_t n_loc, size_t m, int G, int g, int a_rep, fr_t* sL,
                           fr_t* sR, cudaStream_t st) {
  bullet_scalars_kernel<<<grid_for(n_loc), kThreads, 0, st>>>(a, w, n_loc, m, G, g, a_rep, sL, sR);
  LB_LAUNCH_CHECK();
}
// One Bulletproofs round's scalar side in ONE launch (single GPU): fold a, b with the previous round's
// challenge (bullet.rs:127-130), expand the generator weights, form the L / R scalars of the unfolded
// generators in canonical form, and (last CTA, Finalize ticket) the cross inner products c_L, c_R
// (bullet.rs:78-79) + blinds as the two tail columns (Q, h).  Replaces fold_ab + expand_weights + cross_ip +
// reduce + bullet_scalars + set_tail + canonicalize: seven launches of a few microseconds each on the
// critical path of every round.
//   a_in, b_in : length 2m when fold != 0 (folded here into a_out, b_out of length m), else length m
//   w_in       : n/(2m) weights when fold != 0 (expanded into w_out, n/m weights), else n/m weights
//   s_out      : 2 rows x (n/2 + 2) canonical scalars, row 0 = L, row 1 = R, with the generator index of every
//                term in cols_out (same shape): each generator is in exactly one of L, R, so the rows are
//                stored compacted; the last two terms of a row are (c, blind) on the generators n (Q), n+1 (h)
// Thread j <-> generator column j = t*m + pos.  h = m/2:  L gets a'[pos-h] w'[t] G_j for pos >= h (term t*h +
// pos-h), R gets a'[pos+h] w'[t] G_j for pos < h (term t*h + pos).  Threads j < h also own the pair (pos, pos+h)
// of a', b'.
__global__ void __launch_bounds__(kThreads)
    bullet_round_kernel(const fr_t* a_in, const fr_t* b_in, const fr_t* w_in, fr_t* a_out, fr_t* b_out, fr_t* w_out,
                        size_t n, size_t m, int fold, fr_t u, fr_t uinv, fr_t blind_L, fr_t blind_R, fr_t* s_out,
                        uint32_t* cols_out, fr_t* partial, unsigned* counter) {
  __shared__ fr_t scratch[2 * kThreads / 32];
  __shared__ int s_last;
  const size_t h = m / 2, stride = n / 2 + 2;
  const int lg_m = 63 - __clzll((long long)m);
  fr_t acc[2] = {fr_zero(), fr_zero()};
  auto folded_a = [&](size_t i) {
    return fold ? fr_add(fr_mul(ld_fr(a_in + i), u), fr_mul(uinv, ld_fr(a_in + m + i))) : ld_fr(a_in + i);
  };
  auto folded_b = [&](size_t i) {
    return fold ? fr_add(fr_mul(ld_fr(b_in + i), uinv), fr_mul(u, ld_fr(b_in + m + i))) : ld_fr(b_in + i);
  };
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) {
    const size_t t = j >> lg_m, pos = j & (m - 1);  // m is a power of two
    fr_t wt = fold ? fr_mul(ld_fr(w_in + (t >> 1)), (t & 1) ? u : uinv) : ld_fr(w_in + t);
    if (pos == 0 && fold) st_fr(w_out + t, wt);
    const bool is_l = pos >= h;
    const size_t idx = is_l ? pos - h : pos + h;
    const fr_t ai = folded_a(idx);
    const fr_t sc = fr_to_canonical(fr_mul(ai, wt));
    const size_t term = (is_l ? 0 : stride) + t * h + (is_l ? pos - h : pos);
    st_fr(s_out + term, sc);
    cols_out[term] = (uint32_t)j;
    if (t == 0 && pos < h) {  // owner of the pair (pos, pos + h): ai = a'[pos + h]
      const fr_t alo = folded_a(pos), blo = folded_b(pos), bhi = folded_b(idx);
      if (fold) {
        st_fr(a_out + pos, alo);
        st_fr(a_out + idx, ai);
        st_fr(b_out + pos, blo);
        st_fr(b_out + idx, bhi);
      }
      acc[0] = fr_mul(alo, bhi);  // c_L = <a_lo, b_hi>
      acc[1] = fr_mul(ai, blo);   // c_R = <a_hi, b_lo>
    }
  }
  block_sum_fr<2>(acc, scratch);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = acc[0];
    partial[gridDim.x + blockIdx.x] = acc[1];
    __threadfence();
    s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp < 2) {
    fr_t v = fr_zero();
    for (unsigned i = lane; i < gridDim.x; i += 32) v = fr_add(v, ld_fr_cg(partial + (size_t)warp * gridDim.x + i));
    v = warp_sum_fr(v);
    if (lane == 0) {
      st_fr(s_out + (size_t)warp * stride + n / 2, fr_to_canonical(v));
      st_fr(s_out + (size_t)warp * stride + n / 2 + 1, fr_to_canonical(warp == 0 ? blind_L : blind_R));
      cols_out[(size_t)warp * stride + n / 2] = (uint32_t)n;
      cols_out[(size_t)warp * stride + n / 2 + 1] = (uint32_t)(n + 1);
    }
  }
  if (threadIdx.x == 0) *counter = 0;
}
void launch_bullet_round(const fr_t* a_in, const fr_t* b_in, const fr_t* w_in, fr_t* a_out, fr_t* b_out, fr_t* w_out, size_t n,
                         size_t m, int fold, const fr_t& u, const fr_t& uinv, const fr_t& blind_L, const fr_t& blind_R,
                         fr_t* s_out, uint32_t* cols_out, fr_t* partial, unsigned* counter, cudaStream_t st) {
  unsigned blocks = (unsigned)((n + kThreads - 1) / kThreads);
  bullet_round_kernel<<<blocks, kThreads, 0, st>>>(a_in, b_in, w_in, a_out, b_out, w_out, n, m, fold, u, uinv, blind_L,
                                                   blind_R, s_out, cols_out, partial, counter);
  LB_LAUNCH_CHECK();
}
// Two MSM rows over the n + 2 generators (G_0..G_{n-1}, Q, h) in canonical form:
//   row 0 = (k * v[0..n), t00, t01)    row 1 = (0 .. 0, t10, t11)
// i.e. (Cx, Cy) of dot_product.rs:192-197 and (delta, beta) of dot_product.rs:219-230 as ONE two-row MSM.
__global__ void __launch_bounds__(kThreads)
    two_row_scalars_kernel(const fr_t* v, int scale, fr_t k, fr_t t00, fr_t t01, fr_t t10, fr_t t11, size_t n, fr_t* out) {
  const size_t stride = n + 2;
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < stride; j += (size_t)gridDim.x * blockDim.x) {
    fr_t r0, r1 = fr_zero();
    if (j < n) {
      fr_t x = ld_fr(v + j);
      r0 = fr_to_canonical(scale ? fr_mul(x, k) : x);
    } else {
      r0 = fr_to_canonical(j == n ? t00 : t01);
      r1 = fr_to_canonical(j == n ? t10 : t11);
    }
    st_fr(out + j, r0);
    st_fr(out + stride + j, r1);
  }
}
void launch_two_row_scalars(const fr_t* v, int scale, const fr_t& k, const fr_t& t00, const fr_t& t01, const fr_t& t10,
                            const fr_t& t11, size_t n, fr_t* out, cudaStream_t st) {
  two_row_scalars_kernel<<<grid_for(n + 2), kThreads, 0, st>>>(v, scale, k, t00, t01, t10, t11, n, out);
  LB_LAUNCH_CHECK();
}
__global__ void __launch_bounds__(kThreads) scale_kernel(const fr_t* in, fr_t* out, size_t n, fr_t k) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    st_fr(out + i, fr_mul(ld_fr(in + i), k));
}
void launch_scale(const fr_t* in, fr_t* out, size_t n, const fr_t& k, cudaStream_t st) {
  scale_kernel<<<grid_for(n), kThreads, 0, st>>>(in, out, n, k);
  LB_LAUNCH_CHECK();
}

}  // namespace lb
