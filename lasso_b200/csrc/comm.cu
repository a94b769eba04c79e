// lasso_b200 — multi-GPU plumbing for ONE proof sharded over G GPUs: one process per
// GPU, NCCL over NVLink / NVSwitch for the few small exchanges the path has:
//   * per sumcheck round: the (deg+1) or 3*(#circuits) partial sums of every rank -> all ranks
//     (all-gather of the 32-byte Montgomery residues + a modular add on the receiver: NCCL has no
//     "sum mod l" reduction and a limb-wise ncclSum would need widening + carry fix-up anyway);
//   * per row-MSM: the partial (extended-coordinate) points of every rank -> all ranks, added and
//     normalised by the receiver (group addition is not an NCCL op either: "bucket-sum reduce" =
//     gather-then-add);
//   * the log2(G) tail rounds of every sumcheck / Bulletproofs fold, where pairs straddle ranks: the G
//     remaining elements per polynomial are all-gathered once and the tail is computed replicated.
// Every polynomial of global length n >= G is partitioned by the LOW log2(G) index bits: rank g holds
// X[i*G + g].  bound_poly_var_top pairs (i, i + n/2) and both have the same low bits, so every bind and
// every round evaluation is local (src/poly/dense_mlpoly.rs:209-216, src/subprotocols/sumcheck.rs:179-237).
// NCCL is loaded with dlopen so the library has no link-time dependency on it (single-GPU use).
#include <dlfcn.h>
#include <nccl.h>

#include "prover.cuh"

namespace lb {

struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi& nccl() {
  static NcclApi api;
  if (!api.h) {
    const char* names[] = {getenv("LASSO_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n) continue;
      api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.h) break;
    }
    if (!api.h) throw std::runtime_error("cannot dlopen libnccl.so.2 (set LASSO_B200_NCCL)");
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.h, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.h, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.h, "ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))dlsym(api.h, "ncclAllGather");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.h, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather) throw std::runtime_error("libnccl: missing symbols");
  }
  return api;
}
#define LB_NCCL_CHECK(x)                                                                        \
  do {                                                                                          \
    ncclResult_t r_ = (x);                                                                      \
    if (r_ != ncclSuccess)                                                                      \
      throw std::runtime_error(std::string("NCCL error: ") + nccl().GetErrorString(r_));        \
  } while (0)

void comm_unique_id(uint8_t out[128]) {
  ncclUniqueId id;
  LB_NCCL_CHECK(nccl().GetUniqueId(&id));
  static_assert(sizeof(id) == 128, "ncclUniqueId size");
  memcpy(out, &id, 128);
}
void comm_init(Ctx* c, const uint8_t id_bytes[128], int rank, int world) {
  if (world < 1 || (world & (world - 1)) || rank < 0 || rank >= world) throw std::runtime_error("world must be a power of two");
  if (c->nccl_comm) throw std::runtime_error("communicator already initialised");
  c->world = world;
  c->rank = rank;
  c->lg_world = 0;
  while ((1 << c->lg_world) < world) c->lg_world++;
  if (world == 1) return;
  ncclUniqueId id;
  memcpy(&id, id_bytes, 128);
  LB_CUDA_CHECK(cudaSetDevice(c->device));
  ncclComm_t comm;
  LB_NCCL_CHECK(nccl().CommInitRank(&comm, world, id, rank));
  c->nccl_comm = comm;
  c->gather_elems = 1 << 16;
  LB_CUDA_CHECK(cudaMalloc((void**)&c->d_gather, c->gather_elems * sizeof(fr_t)));
}
void comm_destroy(Ctx* c) {
  if (c->nccl_comm) nccl().CommDestroy((ncclComm_t)c->nccl_comm);
  c->nccl_comm = nullptr;
  if (c->d_gather) cudaFree(c->d_gather);
  c->d_gather = nullptr;
}
// recv[g * bytes .. ) = rank g's send buffer, on every rank
void comm_allgather(Ctx* c, const void* d_send, void* d_recv, size_t bytes_per_rank) {
  if (c->world == 1) {
    if (d_send != d_recv) LB_CUDA_CHECK(cudaMemcpyAsync(d_recv, d_send, bytes_per_rank, cudaMemcpyDeviceToDevice, c->st));
    return;
  }
  LB_NCCL_CHECK(nccl().AllGather(d_send, d_recv, bytes_per_rank, ncclChar, (ncclComm_t)c->nccl_comm, c->st));
}

__global__ void sum_gathered_fr_kernel(const fr_t* gathered, int world, int count, fr_t* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  fr_t acc = gathered[i];
  for (int g = 1; g < world; g++) acc = fr_add(acc, gathered[(size_t)g * count + i]);
  out[i] = acc;
}
// d_buf[0..count) <- sum over ranks (mod l), identical on every rank
void comm_allreduce_fr(Ctx* c, fr_t* d_buf, int count) {
  if (c->world == 1) return;
  if ((size_t)count * c->world > c->gather_elems) throw std::runtime_error("allreduce: too many elements");
  comm_allgather(c, d_buf, c->d_gather, (size_t)count * sizeof(fr_t));
  sum_gathered_fr_kernel<<<(count + 127) / 128, 128, 0, c->st>>>(c->d_gather, c->world, count, d_buf);
  g_launches += 1;
}

// tail hand-over: every rank holds ONE element of each of npolys polynomials (local length 1);
// afterwards out[k*G + g] = rank g's element of polynomial k, on every rank (global index = g).
__global__ void pack_heads_kernel(fr_t* const* ptrs, const fr_t* base, size_t stride, int npolys, fr_t* out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= npolys) return;
  out[k] = ptrs ? ptrs[k][0] : base[(size_t)k * stride];
}
__global__ void transpose_gathered_kernel(const fr_t* gathered /*[G][npolys]*/, int world, int npolys, fr_t* out /*[npolys][G]*/) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= world * npolys) return;
  int g = i / npolys, k = i % npolys;
  out[(size_t)k * world + g] = gathered[i];
}
// out[k] = ptrs[k][0] (or base[k*stride]): used to bring the 2*ncirc final claims of a layer to the host in one go
void pack_heads(Ctx* c, fr_t* const* d_ptrs, const fr_t* base, size_t stride, int npolys, fr_t* d_out) {
  pack_heads_kernel<<<(npolys + 127) / 128, 128, 0, c->st>>>(d_ptrs, base, stride, npolys, d_out);
  g_launches += 1;
}
void comm_gather_heads(Ctx* c, fr_t* const* d_ptrs, const fr_t* base, size_t stride, int npolys, fr_t* d_out) {
  if ((size_t)npolys * (c->world + 1) > c->gather_elems) throw std::runtime_error("gather_heads: too many polynomials");
  fr_t* packed = c->d_gather + (size_t)npolys * c->world;
  pack_heads_kernel<<<(npolys + 127) / 128, 128, 0, c->st>>>(d_ptrs, base, stride, npolys, packed);
  comm_allgather(c, packed, c->d_gather, (size_t)npolys * sizeof(fr_t));
  transpose_gathered_kernel<<<(npolys * c->world + 127) / 128, 128, 0, c->st>>>(c->d_gather, c->world, npolys, d_out);
  g_launches += 2;
}

}  // namespace lb
