// lasso_b200 — DensifiedRepresentation::from_lookup_indices on the GPU (src/lasso/densified.rs:33-56;
// SURVEY.md §8f-2).  The reference's timestamp loop is sequential per dimension:
//     ts = final[addr]; read[k] = ts; final[addr] = ts + 1
// i.e. read[k] = #{ j < k : addr[j] == addr[k] } and final[a] = #{ k : addr[k] == a }.  Here, per dimension:
//   1. chunk_hist_kernel : the access sequence is cut into <= 512 chunks; one CTA per chunk histograms its
//                          addresses in a shared-memory table (two 16-bit counters per word) -> P[chunk][addr];
//   2. col_scan_kernel   : one thread per address turns the column P[.][addr] into exclusive prefix counts
//                          (coalesced across addresses) and emits final[addr] = the column total;
//   3. rank_kernel       : one warp per chunk walks its chunk IN ORDER, 32 accesses at a time:
//                          read[k] = P[chunk][addr] + (count of addr so far in this chunk, shared-memory table)
//                                    + (rank among equal addresses inside the warp, __match_any_sync).
// Integer, order-preserving, bit-identical to the sequential scan.  Needs the 2^log_m-entry table in shared
// memory as 16-bit counters: log_m <= 16 (every BASELINE config); larger memories use the host scan.
#include "kernels.cuh"

namespace lb {

static constexpr int kDenseThreads = 1024;

// column `dim` of the row-major n x C index matrix (already narrowed to u32 and range-checked on the host while
// staging it into pinned memory), zero-padded to s (densified.rs:33-37)
__global__ void __launch_bounds__(256)
    extract_dim_kernel(const uint32_t* idx, size_t n, size_t s, int C, int dim, uint32_t* addr) {
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < s; k += (size_t)gridDim.x * blockDim.x)
    addr[k] = k < n ? idx[k * C + dim] : 0u;
}

__global__ void __launch_bounds__(kDenseThreads)
    chunk_hist_kernel(const uint32_t* addr, size_t B, uint32_t m, uint32_t* P) {
  extern __shared__ uint32_t tab[];  // m/2 words, two 16-bit counters each (B <= 32768 so a counter cannot overflow)
  const uint32_t words = (m + 1) / 2;
  for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) tab[w] = 0;
  __syncthreads();
  const uint32_t* a = addr + (size_t)blockIdx.x * B;
  for (size_t k = threadIdx.x; k < B; k += blockDim.x) {
    uint32_t x = a[k];
    atomicAdd(&tab[x >> 1], (x & 1) ? 0x10000u : 1u);
  }
  __syncthreads();
  uint32_t* row = P + (size_t)blockIdx.x * m;
  for (uint32_t x = threadIdx.x; x < m; x += blockDim.x) row[x] = (tab[x >> 1] >> ((x & 1) * 16)) & 0xffffu;
}

// P[c][a] <- sum_{c' < c} P[c'][a];  final_ts (this rank's shard: addresses a = i*G + g) <- column total
__global__ void __launch_bounds__(256)
    col_scan_kernel(uint32_t* P, size_t nchunks, uint32_t m, int G, int g, uint32_t* final_loc) {
  uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= m) return;
  uint32_t run = 0;
  for (size_t c = 0; c < nchunks; c++) {
    uint32_t v = P[c * m + a];
    P[c * m + a] = run;
    run += v;
  }
  if ((int)(a % G) == g) final_loc[a / G] = run;
}

// one warp per chunk; writes this rank's shard of dim / read (access k is local iff k % G == g, at k / G).
// Two different addresses can share a table WORD (x >> 1): their leaders would race on the read-modify-write above.
// Serialise the two halves: even addresses first, then odd ones.
__global__ void __launch_bounds__(32)
    rank_kernel(const uint32_t* addr, size_t B, uint32_t m, const uint32_t* P, int G, int g, uint32_t* dim_loc,
                     uint32_t* read_loc) {
  extern __shared__ uint32_t tab[];
  const uint32_t words = (m + 1) / 2;
  const int lane = threadIdx.x;
  for (uint32_t w = lane; w < words; w += 32) tab[w] = 0;
  __syncwarp();
  const size_t k0 = (size_t)blockIdx.x * B;
  const uint32_t* prow = P + (size_t)blockIdx.x * m;
  for (size_t t = 0; t < B; t += 32) {
    const size_t k = k0 + t + lane;
    const uint32_t x = addr[k];
    const unsigned same = __match_any_sync(0xffffffffu, x);
    const unsigned before = same & ((1u << lane) - 1u);
    const uint32_t in_chunk = (tab[x >> 1] >> ((x & 1) * 16)) & 0xffffu;
    const uint32_t ts = prow[x] + in_chunk + __popc(before);
    __syncwarp();
    if (before == 0 && (x & 1) == 0) tab[x >> 1] += (uint32_t)__popc(same);
    __syncwarp();
    if (before == 0 && (x & 1) == 1) tab[x >> 1] += (uint32_t)__popc(same) << 16;
    __syncwarp();
    if ((int)(k % G) == g) {
      dim_loc[k / G] = x;
      read_loc[k / G] = ts;
    }
  }
}

// function attributes are per device: called from ctx_create for the context's device
void densify_init_device() {
  LB_CUDA_CHECK(cudaFuncSetAttribute(chunk_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  LB_CUDA_CHECK(cudaFuncSetAttribute(rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
}
bool densify_gpu_supported(size_t s, size_t log_m) { return log_m <= 16 && s >= 32; }
size_t densify_chunk(size_t s) {
  size_t B = s / 512;
  if (B < 32) B = 32;
  if (B > 32768) B = 32768;
  return B;
}
// d_idx: n x C u32 on the device.  Scratch: d_addr (s u32), d_P (nchunks * m u32).  Outputs are this rank's shards.
int launch_densify_dim(const uint32_t* d_idx, size_t n, size_t s, int C, int dim, size_t log_m, int G, int g,
                       uint32_t* d_addr, uint32_t* d_P, uint32_t* dim_loc, uint32_t* read_loc, uint32_t* final_loc,
                       cudaStream_t st) {
  const uint32_t m = 1u << log_m;
  const size_t B = densify_chunk(s), nchunks = s / B;
  const size_t smem = (size_t)((m + 1) / 2) * 4;
  size_t eb = (s + 255) / 256;
  if (eb > (size_t)kNumSMs * 8) eb = kNumSMs * 8;
  extract_dim_kernel<<<(unsigned)eb, 256, 0, st>>>(d_idx, n, s, C, dim, d_addr);
  chunk_hist_kernel<<<(unsigned)nchunks, kDenseThreads, smem, st>>>(d_addr, B, m, d_P);
  col_scan_kernel<<<(m + 255) / 256, 256, 0, st>>>(d_P, nchunks, m, G, g, final_loc);
  rank_kernel<<<(unsigned)nchunks, 32, smem, st>>>(d_addr, B, m, d_P, G, g, dim_loc, read_loc);
  LB_LAUNCH_CHECK();
  return 4;
}

}  // namespace lb
