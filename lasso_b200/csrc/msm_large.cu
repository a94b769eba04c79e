// lasso_b200 — K6 for ONE large variable-base MSM (BASELINE config 5: 2^16 .. 2^26 terms): the Pippenger bucket
// method of src/msm/mod.rs:91-164 (msm_bigint_wnaf: signed c-bit digits, one bucket set per window, running-sum
// bucket reduction, c doublings between windows) laid out for a GPU.  The row-batched kernels of msm_kernels.cu are
// built for thousands of short rows over shared generators (fixed c = 8, buckets in shared memory); a single MSM of
// millions of terms wants the reference's large window (c = 13..17: ~16 bucket additions per term instead of 32)
// and buckets in HBM:
//   1. msm_prep_kernel     scalar -> canonical integer (into_bigint, msm/mod.rs:23-25), base -> affine-niels (96 B)
//   2. msm_hist_kernel     signed digits by the offset trick (digit_w = c-bit field of s + sum_w 2^(c-1) 2^(cw), minus
//                          2^(c-1): no carry chain, msm/mod.rs:277-316 yields the same digits), histogram of
//                          (window, |digit|) with RED atomics
//   3. msm_scan_kernel     exclusive scan -> bucket offsets; a bucket with more than S entries (skewed scalars) is
//                          split into units of <= S entries so that no thread walks a long chain
//   4. msm_scatter_kernel  counting-sort scatter of (term | sign) by (window, bucket), one window at a time so that
//                          the window's slice of the list stays in L2
//   5. msm_accum_kernel    one THREAD per unit: a chain of mixed additions over its entries (7 Fq mul each), the
//                          next entry's 96 B in flight during the current addition
//   6. msm_r1_kernel       bucket reduction sum_b b * B_b, level 1: a thread per L consecutive buckets keeps the
//                          running sum / weighted sum of msm/mod.rs:139-145 locally
//      msm_r2_kernel       level 2: one CTA per window combines the L-blocks with a tree that carries
//                          (sum, index-weighted sum): w = w_l + w_r + h * sum_r, h = 2^k by k doublings; then the
//                          window's weight 2^(c w) by c*w doublings (msm/mod.rs:150-163 does the same c doublings per
//                          window, serially over the windows)
//   7. msm_final_kernel    adds the window totals, normalises.
// Integer-ALU bound: reported as mixed additions/s against the 7.2 G/s the row-commitment kernel reaches.
// Same group element as msm_bigint_wnaf for every input; outputs are compared after affine normalisation.
#if defined(__CUDACC__)
#define LB_FQ_MUL_ATTR static __host__ __device__ __noinline__
#define LB_FQ_MUL_BYVALUE
#endif
#include "kernels.cuh"
#include "msm.cuh"

namespace lb {

namespace {

__device__ __forceinline__ pt_niels ldn(const pt_niels* p) {
  pt_niels n;
  n.yplusx = ld_fq(&p->yplusx);
  n.yminusx = ld_fq(&p->yminusx);
  n.t2d = ld_fq(&p->t2d);
  return n;
}
__device__ __forceinline__ void stn(pt_niels* p, const pt_niels& n) {
  st_fq(&p->yplusx, n.yplusx);
  st_fq(&p->yminusx, n.yminusx);
  st_fq(&p->t2d, n.t2d);
}
__device__ __forceinline__ pt_ext ldp(const pt_ext* p) {
  pt_ext r;
  r.X = ld_fq(&p->X);
  r.Y = ld_fq(&p->Y);
  r.Z = ld_fq(&p->Z);
  r.T = ld_fq(&p->T);
  return r;
}
__device__ __forceinline__ void stp(pt_ext* p, const pt_ext& r) {
  st_fq(&p->X, r.X);
  st_fq(&p->Y, r.Y);
  st_fq(&p->Z, r.Z);
  st_fq(&p->T, r.T);
}

// biased scalar s + sum_{w < nw} 2^(c-1) * 2^(c w) as 9 x u32 (c * nw <= 272 bits); digit w = its c-bit field w
// minus 2^(c-1), in [-2^(c-1), 2^(c-1))
struct Biased {
  uint32_t b[9];
  __device__ __forceinline__ Biased(const uint32_t s[8], const uint32_t off[9]) {
    uint32_t carry = 0;
#pragma unroll
    for (int l = 0; l < 9; l++) {
      const uint64_t t = (uint64_t)(l < 8 ? s[l] : 0u) + off[l] + carry;
      b[l] = (uint32_t)t;
      carry = (uint32_t)(t >> 32);
    }
  }
  __device__ __forceinline__ int digit(int w, int c) const {
    const int bit = w * c, limb = bit >> 5, sh = bit & 31;
    uint64_t two = 0;
#pragma unroll
    for (int l = 0; l < 9; l++) {  // b[limb] | b[limb+1] << 32 without dynamic register indexing
      if (l == limb) two |= b[l];
      if (l == limb + 1) two |= (uint64_t)b[l] << 32;
    }
    return (int)((two >> sh) & ((1u << c) - 1u)) - (1 << (c - 1));
  }
};
struct BiasOff {
  uint32_t v[9];
};

}  // namespace

// ---------------------------------------------------------------------------------------------- 1. prep
__global__ void __launch_bounds__(256)
    msm_prep_kernel(const fq_t* bases_ark, const fr_t* scalars_mont, size_t n, size_t n_pool, pt_niels* niels, fr_t* canon,
                    unsigned* max_bits) {
  unsigned mb = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t j = n_pool ? i % n_pool : i;  // bench inputs: a pool of distinct points tiled to n terms
    stn(niels + i, niels_from_ark_affine(ld_fq(bases_ark + 2 * j), ld_fq(bases_ark + 2 * j + 1)));
    const fr_t c = fr_to_canonical(ld_fr(scalars_mont + i));
    st_fr(canon + i, c);
    unsigned b = 0;
#pragma unroll
    for (int l = 0; l < 8; l++)
      if (c.v[l]) b = 32 * l + (32 - __clz(c.v[l]));
    mb = b > mb ? b : mb;
  }
  mb = __reduce_max_sync(0xffffffffu, mb);
  if ((threadIdx.x & 31) == 0 && mb) atomicMax(max_bits, mb);
}

// ---------------------------------------------------------------------------------------------- 2. histogram
__global__ void __launch_bounds__(256)
    msm_hist_kernel(const fr_t* canon, size_t n, int c, int nw, uint32_t NB1, BiasOff off, uint32_t* cnt) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const fr_t s = ld_fr(canon + i);
    const Biased bs(s.v, off.v);
    for (int w = 0; w < nw; w++) {
      const int d = bs.digit(w, c);
      if (d) atomicAdd(cnt + (size_t)w * NB1 + (uint32_t)(d < 0 ? -d : d), 1u);
    }
  }
}

// ---------------------------------------------------------------------------------------------- 3. scan
// off[i] = sum_{j < i} cnt[j], uoff[i] = sum_{j < i} units(cnt[j]) with units(x) = ceil(x / S); one CTA.
// totals[0] = number of entries, totals[1] = number of units.
__global__ void __launch_bounds__(1024)
    msm_scan_kernel(const uint32_t* cnt, uint32_t total, uint32_t S, uint32_t* off, uint32_t* uoff, uint32_t* totals) {
  __shared__ uint32_t s1[1024], s2[1024];
  const uint32_t t = threadIdx.x, chunk = (total + 1023u) / 1024u;
  const uint32_t lo = min(total, t * chunk), hi = min(total, lo + chunk);
  uint32_t a = 0, b = 0;
  for (uint32_t i = lo; i < hi; i++) {
    const uint32_t x = cnt[i];
    a += x;
    b += (x + S - 1) / S;
  }
  s1[t] = a;
  s2[t] = b;
  __syncthreads();
  for (uint32_t d = 1; d < 1024; d <<= 1) {  // inclusive Hillis-Steele scan
    uint32_t x1 = 0, x2 = 0;
    if (t >= d) {
      x1 = s1[t - d];
      x2 = s2[t - d];
    }
    __syncthreads();
    s1[t] += x1;
    s2[t] += x2;
    __syncthreads();
  }
  uint32_t r1 = s1[t] - a, r2 = s2[t] - b;
  for (uint32_t i = lo; i < hi; i++) {
    const uint32_t x = cnt[i];
    off[i] = r1;
    uoff[i] = r2;
    r1 += x;
    r2 += (x + S - 1) / S;
  }
  if (t == 1023) {
    off[total] = s1[1023];
    uoff[total] = s2[1023];
    totals[0] = s1[1023];
    totals[1] = s2[1023];
  }
}
// unit -> bucket map
__global__ void __launch_bounds__(256)
    msm_unit_map_kernel(const uint32_t* uoff, uint32_t total, uint32_t* unit_bucket) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const uint32_t u0 = uoff[i], u1 = uoff[i + 1];
  for (uint32_t u = u0; u < u1; u++) unit_bucket[u] = i;
}

// ---------------------------------------------------------------------------------------------- 4. scatter
// blockIdx.y = window: the CTAs of one window run together, its slice of `entries` (n x 4 B) stays in L2
__global__ void __launch_bounds__(256)
    msm_scatter_kernel(const fr_t* canon, size_t n, int c, uint32_t NB1, BiasOff off, const uint32_t* boff, uint32_t* fill,
                       uint32_t* entries) {
  const int w = blockIdx.y;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const fr_t s = ld_fr(canon + i);
    const Biased bs(s.v, off.v);
    const int d = bs.digit(w, c);
    if (d) {
      const size_t bi = (size_t)w * NB1 + (uint32_t)(d < 0 ? -d : d);
      const uint32_t pos = boff[bi] + atomicAdd(fill + bi, 1u);
      entries[pos] = (uint32_t)i | (d < 0 ? 0x80000000u : 0u);
    }
  }
}

// ---------------------------------------------------------------------------------------------- 5. accumulate
__global__ void __launch_bounds__(128)
    msm_accum_kernel(const pt_niels* niels, const uint32_t* entries, const uint32_t* cnt, const uint32_t* boff,
                     const uint32_t* uoff, const uint32_t* unit_bucket, const uint32_t* totals, uint32_t S, pt_ext* unit_sum) {
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= totals[1]) return;
  const uint32_t bi = unit_bucket[u], k = u - uoff[bi];
  const uint32_t lo = boff[bi] + k * S, end = boff[bi] + cnt[bi], hi = min(end, lo + S);
  pt_ext acc = pt_identity();
  uint32_t e = entries[lo];
  pt_niels nn = ldn(niels + (e & 0x7fffffffu));
  for (uint32_t p = lo; p < hi; p++) {
    const uint32_t ecur = e;
    const pt_niels ncur = nn;
    if (p + 1 < hi) {
      e = entries[p + 1];
      nn = ldn(niels + (e & 0x7fffffffu));
    }
    acc = pt_madd(acc, (ecur & 0x80000000u) ? niels_neg(ncur) : ncur);
  }
  stp(unit_sum + u, acc);
}

// ---------------------------------------------------------------------------------------------- 6. bucket reduction
// A bucket that was split into several units (skewed scalars: e.g. the top window of 20-bit scalars has a handful of
// non-empty buckets holding all n terms) is combined here, one WARP per bucket: lanes stride over the units, then a
// shuffle tree; the sum replaces the bucket's first unit.  Buckets with <= 1 unit (all of them for uniform scalars)
// leave immediately.
__global__ void __launch_bounds__(256)
    msm_unit_combine_kernel(pt_ext* unit_sum, const uint32_t* uoff, uint32_t total) {
  const uint32_t bi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (bi >= total) return;
  const uint32_t u0 = uoff[bi], u1 = uoff[bi + 1];
  if (u1 - u0 <= 1) return;  // warp-uniform
  pt_ext acc = pt_identity();
  bool any = false;
  for (uint32_t u = u0 + lane; u < u1; u += 32) {
    const pt_ext p = ldp(unit_sum + u);
    acc = any ? pt_add(acc, p) : p;
    any = true;
  }
#pragma unroll 1
  for (int d = 16; d >= 1; d >>= 1) {
    pt_ext o;
#pragma unroll
    for (int l = 0; l < 8; l++) {
      o.X.v[l] = __shfl_down_sync(0xffffffffu, acc.X.v[l], d);
      o.Y.v[l] = __shfl_down_sync(0xffffffffu, acc.Y.v[l], d);
      o.Z.v[l] = __shfl_down_sync(0xffffffffu, acc.Z.v[l], d);
      o.T.v[l] = __shfl_down_sync(0xffffffffu, acc.T.v[l], d);
    }
    acc = pt_add(acc, o);
  }
  if (lane == 0) stp(unit_sum + u0, acc);
}
// B_b of (window w, bucket b): its first unit (the whole bucket after msm_unit_combine_kernel)
__device__ __forceinline__ bool bucket_sum(const pt_ext* unit_sum, const uint32_t* uoff, size_t bi, pt_ext& out) {
  const uint32_t u0 = uoff[bi], u1 = uoff[bi + 1];
  if (u0 == u1) return false;
  out = ldp(unit_sum + u0);
  return true;
}
// thread (w, t): buckets t*L + 1 .. t*L + L.  acc = sum_j j * B_{tL+j}, run = sum_j B_{tL+j}
__global__ void __launch_bounds__(128)
    msm_r1_kernel(const pt_ext* unit_sum, const uint32_t* uoff, int nw, uint32_t NB1, uint32_t L, uint32_t T2, pt_ext* r1_acc,
                  pt_ext* r1_run) {
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (uint32_t)nw * T2) return;
  const uint32_t w = id / T2, t = id - w * T2;
  pt_ext run = pt_identity(), acc = pt_identity();
  bool any = false;
  for (uint32_t j = L; j >= 1; j--) {
    pt_ext B;
    if (bucket_sum(unit_sum, uoff, (size_t)w * NB1 + (size_t)t * L + j, B)) {
      run = any ? pt_add(run, B) : B;
      any = true;
    }
    if (any) acc = pt_add(acc, run);
  }
  stp(r1_acc + id, acc);
  stp(r1_run + id, run);
}
// shared-memory point storage (SoA): element (coord c, limb l) of point idx at base[(c*8 + l) * n + idx]
__device__ __forceinline__ void sm_st(uint32_t* base, int n, int idx, const pt_ext& p) {
#pragma unroll
  for (int l = 0; l < 8; l++) {
    base[(0 * 8 + l) * n + idx] = p.X.v[l];
    base[(1 * 8 + l) * n + idx] = p.Y.v[l];
    base[(2 * 8 + l) * n + idx] = p.Z.v[l];
    base[(3 * 8 + l) * n + idx] = p.T.v[l];
  }
}
__device__ __forceinline__ pt_ext sm_ld(const uint32_t* base, int n, int idx) {
  pt_ext p;
#pragma unroll
  for (int l = 0; l < 8; l++) {
    p.X.v[l] = base[(0 * 8 + l) * n + idx];
    p.Y.v[l] = base[(1 * 8 + l) * n + idx];
    p.Z.v[l] = base[(2 * 8 + l) * n + idx];
    p.T.v[l] = base[(3 * 8 + l) * n + idx];
  }
  return p;
}
// one CTA per window, T2 threads (a power of two <= 512).  Tree over the L-blocks carrying
//   x = sum of acc, s = sum of run, y = sum_t t * run_t (index inside the current subtree):
//   merging [left | right] of h blocks each: y = y_l + y_r + h * s_r.
// window total = x + L * y, times 2^(c w).
__global__ void __launch_bounds__(512)
    msm_r2_kernel(const pt_ext* r1_acc, const pt_ext* r1_run, uint32_t T2, uint32_t lgL, int c, pt_ext* win_total) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* sx = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* ss = sx + 32 * T2;
  uint32_t* sy = ss + 32 * T2;
  const int w = blockIdx.x, t = threadIdx.x, n = (int)T2;
  pt_ext x = ldp(r1_acc + (size_t)w * T2 + t), s = ldp(r1_run + (size_t)w * T2 + t), y = pt_identity();
  sm_st(sx, n, t, x);
  sm_st(ss, n, t, s);
  sm_st(sy, n, t, y);
  __syncthreads();
  int lg = 0;
  for (uint32_t h = 1; h < T2; h <<= 1, lg++) {
    const bool act = (t & (2 * h - 1)) == 0;
    if (act) {
      const pt_ext xr = sm_ld(sx, n, t + h), sr = sm_ld(ss, n, t + h), yr = sm_ld(sy, n, t + h);
      pt_ext hs = sr;
      for (int k = 0; k < lg; k++) hs = pt_dbl(hs);  // h * s_r
      x = pt_add(x, xr);
      y = pt_add(pt_add(y, yr), hs);
      s = pt_add(s, sr);
    }
    __syncthreads();
    if (act) {
      sm_st(sx, n, t, x);
      sm_st(ss, n, t, s);
      sm_st(sy, n, t, y);
    }
    __syncthreads();
  }
  if (t == 0) {
    for (uint32_t k = 0; k < lgL; k++) y = pt_dbl(y);  // L * y
    pt_ext tot = pt_add(x, y);
    for (int k = 0; k < c * w; k++) tot = pt_dbl(tot);  // 2^(c w)
    stp(win_total + w, tot);
  }
}
// ---------------------------------------------------------------------------------------------- 7. final
__global__ void __launch_bounds__(32)
    msm_final_kernel(const pt_ext* win_total, int nw, fq_t* out_ext, uint32_t* out_raw) {
  const int lane = threadIdx.x;
  pt_ext acc = lane < nw ? ldp(win_total + lane) : pt_identity();
#pragma unroll 1
  for (int d = 16; d >= 1; d >>= 1) {
    pt_ext o;
#pragma unroll
    for (int l = 0; l < 8; l++) {
      o.X.v[l] = __shfl_down_sync(0xffffffffu, acc.X.v[l], d);
      o.Y.v[l] = __shfl_down_sync(0xffffffffu, acc.Y.v[l], d);
      o.Z.v[l] = __shfl_down_sync(0xffffffffu, acc.Z.v[l], d);
      o.T.v[l] = __shfl_down_sync(0xffffffffu, acc.T.v[l], d);
    }
    acc = pt_add(acc, o);
  }
  if (lane == 0) {
    if (out_raw) {
#pragma unroll
      for (int l = 0; l < 8; l++) {
        out_raw[l] = acc.X.v[l];
        out_raw[8 + l] = acc.Y.v[l];
        out_raw[16 + l] = acc.Z.v[l];
        out_raw[24 + l] = acc.T.v[l];
      }
    }
    if (out_ext) {
      fq_t x, y;
      pt_to_affine_canonical(acc, x, y);
      out_ext[0] = fq_to_ark(x);
      out_ext[1] = fq_to_ark(y);
      out_ext[2] = fq_to_ark(fq_mul(x, y));
      out_ext[3] = fq_to_ark(fq_one());
    }
  }
}

// ---------------------------------------------------------------------------------------------- naive cross-check
// An independent evaluation of the same sum for the parity tests at sizes the CPU oracle cannot reach: every term
// by plain double-and-add over the bits of its canonical scalar (no digits, no buckets, no tables), then a tree sum.
__global__ void __launch_bounds__(128)
    msm_naive_terms_kernel(const fq_t* bases_ark, const fr_t* scalars_mont, size_t n, size_t n_pool, pt_ext* partial) {
  __shared__ uint32_t buf[32 * 128];
  pt_ext acc = pt_identity();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t j = n_pool ? i % n_pool : i;
    const pt_niels b = niels_from_ark_affine(ld_fq(bases_ark + 2 * j), ld_fq(bases_ark + 2 * j + 1));
    const fr_t s = fr_to_canonical(ld_fr(scalars_mont + i));
    pt_ext t = pt_identity();
    int top = -1;
    for (int l = 7; l >= 0 && top < 0; l--)
      if (s.v[l]) top = 32 * l + 31 - __clz(s.v[l]);
    for (int bit = top; bit >= 0; bit--) {
      t = pt_dbl(t);
      if ((s.v[bit >> 5] >> (bit & 31)) & 1u) t = pt_madd(t, b);
    }
    acc = pt_add(acc, t);
  }
  sm_st(buf, 128, threadIdx.x, acc);
  __syncthreads();
  for (int d = 64; d >= 1; d >>= 1) {
    if ((int)threadIdx.x < d) {
      acc = pt_add(acc, sm_ld(buf, 128, threadIdx.x + d));
      sm_st(buf, 128, threadIdx.x, acc);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) stp(partial + blockIdx.x, acc);
}
__global__ void __launch_bounds__(128) msm_naive_sum_kernel(const pt_ext* partial, int count, fq_t* out_ext) {
  __shared__ uint32_t buf[32 * 128];
  pt_ext acc = pt_identity();
  for (int i = threadIdx.x; i < count; i += blockDim.x) acc = pt_add(acc, ldp(partial + i));
  sm_st(buf, 128, threadIdx.x, acc);
  __syncthreads();
  for (int d = 64; d >= 1; d >>= 1) {
    if ((int)threadIdx.x < d) {
      acc = pt_add(acc, sm_ld(buf, 128, threadIdx.x + d));
      sm_st(buf, 128, threadIdx.x, acc);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    fq_t x, y;
    pt_to_affine_canonical(acc, x, y);
    out_ext[0] = fq_to_ark(x);
    out_ext[1] = fq_to_ark(y);
    out_ext[2] = fq_to_ark(fq_mul(x, y));
    out_ext[3] = fq_to_ark(fq_one());
  }
}
void launch_msm_naive(const fq_t* bases_ark, const fr_t* scalars_mont, size_t n, size_t n_pool, pt_ext* partial /* 1184 */,
                      fq_t* out_ext, cudaStream_t st) {
  const int blocks = kNumSMs * 8;
  msm_naive_terms_kernel<<<blocks, 128, 0, st>>>(bases_ark, scalars_mont, n, n_pool, partial);
  LB_LAUNCH_CHECK();
  msm_naive_sum_kernel<<<1, 128, 0, st>>>(partial, blocks, out_ext);
  LB_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------- host side
int msm_large_window_bits(size_t n) {
  // the reference's rule is c = floor(log2(n) * 0.69) + 2 (msm/mod.rs:112-116, 322-325): 13 at 2^16, 17 at 2^22, 19 at 2^26;
  // here the bucket reduction's serial chains cap it at 17
  int lg = 0;
  while (((size_t)1 << (lg + 1)) <= n) lg++;
  int c = (int)(lg * 0.69) + 2;
  if (c < 8) c = 8;
  if (c > 17) c = 17;
  return c;
}
MsmLargePlan msm_large_plan(size_t n, unsigned max_bits) {
  MsmLargePlan p;
  p.n = n;
  p.nbits = max_bits < 1 ? 1 : (int)max_bits;
  p.c = msm_large_window_bits(n);
  if (p.c > p.nbits + 1) p.c = p.nbits + 1 < 2 ? 2 : p.nbits + 1;  // small scalars: one window holds them
  p.nw = (p.nbits + 2 + p.c - 1) / p.c;                             // c nw >= nbits + 2  =>  s + bias < 2^(c nw)
  if (p.nw > 32) throw std::runtime_error("msm_large: more than 32 windows");
  if (p.c * p.nw > 9 * 32 - 1) throw std::runtime_error("msm_large: biased scalar wider than 9 limbs");
  p.NB = 1u << (p.c - 1);
  p.NB1 = p.NB + 1;
  p.T2 = p.NB < 512 ? p.NB : 512;
  p.L = p.NB / p.T2;
  p.lgL = 0;
  while ((1u << p.lgL) < p.L) p.lgL++;
  const size_t avg = (n + p.NB - 1) / p.NB;
  p.S = (uint32_t)std::max<size_t>(64, 4 * avg);
  p.total = (uint32_t)p.nw * p.NB1;
  p.max_entries = n * (size_t)p.nw;
  if (p.max_entries >= ((size_t)1 << 32) || n >= ((size_t)1 << 31)) throw std::runtime_error("msm_large: too many terms");
  p.max_units = (size_t)p.total + p.max_entries / p.S + 1;
  return p;
}
size_t msm_large_scratch_bytes(const MsmLargePlan& p) {
  size_t b = 0;
  b += 3 * ((size_t)p.total + 1) * 4 + 16;  // cnt/fill, off, uoff (+ totals)
  b += (size_t)p.total * 4;
  b += p.max_entries * 4;                    // entries
  b += p.max_units * 4;                      // unit_bucket
  b += p.max_units * sizeof(pt_ext);         // unit_sum
  b += 2 * (size_t)p.nw * p.T2 * sizeof(pt_ext) + 64 * sizeof(pt_ext);
  return b + 4096;
}
void msm_large_init_device() {
  LB_CUDA_CHECK(cudaFuncSetAttribute(msm_r2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 32 * 512 * 4));
}
void launch_msm_large_prep(const fq_t* bases_ark, const fr_t* scalars_mont, size_t n, size_t n_pool, pt_niels* niels,
                           fr_t* canon, unsigned* d_max_bits, cudaStream_t st) {
  size_t b = (n + 255) / 256;
  if (b > (size_t)kNumSMs * 8) b = kNumSMs * 8;
  msm_prep_kernel<<<(unsigned)b, 256, 0, st>>>(bases_ark, scalars_mont, n, n_pool, niels, canon, d_max_bits);
  LB_LAUNCH_CHECK();
}
// scratch: msm_large_scratch_bytes(plan) bytes.  Outputs (either may be null): out_ext = (x, y, t, z = 1) arkworks
// limbs; out_raw = un-normalised (X, Y, Z, T) internal limbs (for the cross-GPU gather-then-add).  Returns the number
// of kernels launched.
int launch_msm_large(const MsmLargePlan& p, const pt_niels* niels, const fr_t* canon, void* scratch, fq_t* out_ext,
                     uint32_t* out_raw, cudaStream_t st) {
  uint8_t* s = (uint8_t*)scratch;
  auto take = [&](size_t bytes) {
    uint8_t* r = s;
    s += (bytes + 255) & ~(size_t)255;
    return r;
  };
  uint32_t* cnt = (uint32_t*)take(((size_t)p.total + 1) * 4);
  uint32_t* fill = (uint32_t*)take(((size_t)p.total + 1) * 4);
  uint32_t* off = (uint32_t*)take(((size_t)p.total + 1) * 4);
  uint32_t* uoff = (uint32_t*)take(((size_t)p.total + 1) * 4);
  uint32_t* totals = (uint32_t*)take(16);
  uint32_t* entries = (uint32_t*)take(p.max_entries * 4);
  uint32_t* unit_bucket = (uint32_t*)take(p.max_units * 4);
  pt_ext* unit_sum = (pt_ext*)take(p.max_units * sizeof(pt_ext));
  pt_ext* r1_acc = (pt_ext*)take((size_t)p.nw * p.T2 * sizeof(pt_ext));
  pt_ext* r1_run = (pt_ext*)take((size_t)p.nw * p.T2 * sizeof(pt_ext));
  pt_ext* win_total = (pt_ext*)take(64 * sizeof(pt_ext));
  BiasOff bo;  // sum_{w < nw} 2^(c-1) * 2^(c w)
  for (int l = 0; l < 9; l++) bo.v[l] = 0;
  for (int w = 0; w < p.nw; w++) {
    const int bit = w * p.c + p.c - 1;
    bo.v[bit >> 5] |= 1u << (bit & 31);
  }
  LB_CUDA_CHECK(cudaMemsetAsync(cnt, 0, ((size_t)p.total + 1) * 4, st));
  LB_CUDA_CHECK(cudaMemsetAsync(fill, 0, ((size_t)p.total + 1) * 4, st));
  size_t b = (p.n + 255) / 256;
  if (b > (size_t)kNumSMs * 8) b = kNumSMs * 8;
  msm_hist_kernel<<<(unsigned)b, 256, 0, st>>>(canon, p.n, p.c, p.nw, p.NB1, bo, cnt);
  LB_LAUNCH_CHECK();
  msm_scan_kernel<<<1, 1024, 0, st>>>(cnt, p.total, p.S, off, uoff, totals);
  LB_LAUNCH_CHECK();
  msm_unit_map_kernel<<<(p.total + 255) / 256, 256, 0, st>>>(uoff, p.total, unit_bucket);
  LB_LAUNCH_CHECK();
  {
    size_t bx = (p.n + 255) / 256;
    if (bx > (size_t)kNumSMs * 4) bx = kNumSMs * 4;
    dim3 grid((unsigned)bx, (unsigned)p.nw);
    msm_scatter_kernel<<<grid, 256, 0, st>>>(canon, p.n, p.c, p.NB1, bo, off, fill, entries);
    LB_LAUNCH_CHECK();
  }
  msm_accum_kernel<<<(unsigned)((p.max_units + 127) / 128), 128, 0, st>>>(niels, entries, cnt, off, uoff, unit_bucket, totals, p.S,
                                                                        unit_sum);
  LB_LAUNCH_CHECK();
  msm_unit_combine_kernel<<<(unsigned)(((size_t)p.total * 32 + 255) / 256), 256, 0, st>>>(unit_sum, uoff, p.total);
  LB_LAUNCH_CHECK();
  msm_r1_kernel<<<(unsigned)(((size_t)p.nw * p.T2 + 127) / 128), 128, 0, st>>>(unit_sum, uoff, p.nw, p.NB1, p.L, p.T2, r1_acc,
                                                                               r1_run);
  LB_LAUNCH_CHECK();
  msm_r2_kernel<<<p.nw, p.T2, 3 * 32 * (size_t)p.T2 * 4, st>>>(r1_acc, r1_run, p.T2, p.lgL, p.c, win_total);
  LB_LAUNCH_CHECK();
  msm_final_kernel<<<1, 32, 0, st>>>(win_total, p.nw, out_ext, out_raw);
  LB_LAUNCH_CHECK();
  return 9;
}

}  // namespace lb
