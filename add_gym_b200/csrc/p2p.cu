// The one exchange step of the path as ONE kernel over NVLink peer memory: gradient reduce-scatter -> AdamW on the
// shard this rank owns -> all-gather of the updated parameters, with the cross-GPU synchronisation done by flags in peer
// memory.  Replaces  dist.all_reduce(flat_grad) + addk_adamw  (NCCL all-reduce of 17.4 MB + a separate optimizer launch
// per optimizer step; the reference's DDP wrapper would have issued the same all-reduce).
//
// Every rank owns, in memory the peers can address (cudaIpc handles exchanged once):
//   grad  [P]     its local flat gradient (written by the slab reduction of addk_update_minibatch)
//   param [P]     the flat parameter vector (peers write the shards they updated into it)
//   flags [2][8]  uint32 epochs: flags[0][p] = "peer p's gradient of epoch e is complete",
//                                flags[1][p] = "peer p has written its shard of epoch e into MY param"
// Per optimizer step (epoch = optimizer step number, the same on all ranks):
//   0. announce my gradient (remote store of the epoch into every peer's flags[0][me]); wait for every peer's;
//   1. for the shard [me * chunk, (me + 1) * chunk): g = sum over the ranks in rank order of grad_p (remote 128-bit
//      loads), AdamW with g / world on my exp_avg / exp_avg_sq shard, new parameters stored into EVERY rank's param;
//   2. fence, then the last CTA announces "my shard is everywhere" to every peer and waits until every peer's shard has
//      arrived here -- so the kernel only completes when this rank's parameter vector is whole again.
// All ranks hold bit-identical parameters afterwards (each value is computed once and copied).  Data volume per rank:
// (world-1)/world * 17.4 MB read + the same written over NVLink (8 GPUs: 15 + 15 MB ~ 35 us at 900 GB/s).
// Moments are sharded: only the owner's exp_avg / exp_avg_sq shard is current (state_dict gathers them).
#include "common.cuh"
#include "addk.h"
#include "adam.cuh"
#include <string.h>

namespace addk {

constexpr int P2P_MAX_RANKS = 8;
struct P2PPeers { float* grad[P2P_MAX_RANKS]; float* param[P2P_MAX_RANKS]; unsigned int* flags[P2P_MAX_RANKS]; };

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {      // peer memory: never from a stale cache line
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void wait_epoch(const unsigned int* flag, unsigned int epoch) {
  if ((int)(ld_acquire_sys(flag) - epoch) >= 0) return;
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
    __nanosleep(64);
    if (clock64() - t0 > 20000000000LL) __trap();      // ~10 s: a rank died -- fail instead of hanging the device
  }
}

__global__ void __launch_bounds__(256) p2p_adamw_kernel(const P2PPeers peers, int rank, int world, float* __restrict__ m,
                                                        float* __restrict__ v, long long n, unsigned int epoch, const AdamK k,
                                                        unsigned int* __restrict__ ticket) {
  __shared__ unsigned int s_last;
  unsigned int* const my_flags = peers.flags[rank];
  // ---- 0. gradients ready everywhere
  if (blockIdx.x == 0 && (int)threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(peers.flags[threadIdx.x] + rank, epoch);
  }
  if ((int)threadIdx.x < world) wait_epoch(my_flags + threadIdx.x, epoch);
  __syncthreads();
  // ---- 1. my shard: reduce, AdamW, broadcast
  const long long n4 = (n + 3) >> 2;                                  // float4 slots (the vectors are padded to 4)
  const long long per = (n4 + world - 1) / world;
  const long long q0 = (long long)rank * per, q1 = (q0 + per < n4) ? q0 + per : n4;
  for (long long q = q0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; q < q1; q += (long long)gridDim.x * blockDim.x) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < world; ++p) {                                // fixed order: every rank would compute the same bits
      const float4 t = (p == rank) ? *reinterpret_cast<const float4*>(peers.grad[p] + 4 * q) : ld_peer4(peers.grad[p] + 4 * q);
      g.x = add_rn(g.x, t.x); g.y = add_rn(g.y, t.y); g.z = add_rn(g.z, t.z); g.w = add_rn(g.w, t.w);
    }
    float4 P = *reinterpret_cast<const float4*>(peers.param[rank] + 4 * q), M = *reinterpret_cast<const float4*>(m + 4 * q),
           V = *reinterpret_cast<const float4*>(v + 4 * q);
    adam1(k, P.x, g.x, M.x, V.x); adam1(k, P.y, g.y, M.y, V.y); adam1(k, P.z, g.z, M.z, V.z); adam1(k, P.w, g.w, M.w, V.w);
    *reinterpret_cast<float4*>(m + 4 * q) = M;
    *reinterpret_cast<float4*>(v + 4 * q) = V;
    for (int p = 0; p < world; ++p) *reinterpret_cast<float4*>(peers.param[p] + 4 * q) = P;
  }
  // ---- 2. my shard is everywhere; wait until every peer's shard is here
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if ((int)threadIdx.x < world) {
    st_release_sys(peers.flags[threadIdx.x] + P2P_MAX_RANKS + rank, epoch);
    wait_epoch(my_flags + P2P_MAX_RANKS + threadIdx.x, epoch);
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

}  // namespace addk

using namespace addk;

extern "C" int addk_p2p_alloc(long long bytes, void** ptr_out) {
  if (!ptr_out || bytes <= 0) return ADDK_ERR_ARG;
  void* p = nullptr;
  if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess || cudaMemset(p, 0, (size_t)bytes) != cudaSuccess) {
    addk_set_error("p2p: cudaMalloc failed"); return ADDK_ERR_LAUNCH;
  }
  *ptr_out = p;
  return ADDK_OK;
}
extern "C" int addk_p2p_free(void* ptr) { return cudaFree(ptr) == cudaSuccess ? ADDK_OK : ADDK_ERR_LAUNCH; }
extern "C" int addk_p2p_export(void* ptr, unsigned char* handle64_host) {
  if (!ptr || !handle64_host) return ADDK_ERR_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, ptr) != cudaSuccess) { addk_set_error(cudaGetErrorString(cudaGetLastError())); return ADDK_ERR_LAUNCH; }
  memcpy(handle64_host, &h, 64);
  return ADDK_OK;
}
extern "C" int addk_p2p_open(const unsigned char* handle64_host, void** ptr_out) {
  if (!handle64_host || !ptr_out) return ADDK_ERR_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64_host, 64);
  void* p = nullptr;
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
    addk_set_error(cudaGetErrorString(cudaGetLastError())); return ADDK_ERR_LAUNCH;
  }
  *ptr_out = p;
  return ADDK_OK;
}

extern "C" int addk_p2p_adamw(void* stream, int rank, int world, float* const* grad_ptrs_host, float* const* param_ptrs_host,
                              unsigned int* const* flag_ptrs_host, float* exp_avg, float* exp_avg_sq, long long n,
                              int step, double lr, double beta1, double beta2, double eps, double weight_decay,
                              unsigned int* ticket, int max_blocks) {
  if (!grad_ptrs_host || !param_ptrs_host || !flag_ptrs_host || !exp_avg || !exp_avg_sq || !ticket || n <= 0 || step < 1 ||
      world < 1 || world > P2P_MAX_RANKS || rank < 0 || rank >= world)
    return ADDK_ERR_ARG;
  P2PPeers peers;
  for (int p = 0; p < P2P_MAX_RANKS; ++p) {
    peers.grad[p] = p < world ? grad_ptrs_host[p] : nullptr;
    peers.param[p] = p < world ? param_ptrs_host[p] : nullptr;
    peers.flags[p] = p < world ? flag_ptrs_host[p] : nullptr;
    if (p < world && (!peers.grad[p] || !peers.param[p] || !peers.flags[p] || ((uintptr_t)peers.grad[p] & 15) || ((uintptr_t)peers.param[p] & 15)))
      return ADDK_ERR_ARG;
  }
  if (((uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) return ADDK_ERR_ARG;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  const AdamK k = {(float)(1.0 - lr * weight_decay), (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)(lr / bc1),
                   (float)sqrt(bc2), (float)eps, (float)(1.0 / world)};
  // every CTA waits inside the kernel: the grid must be co-resident (one wave)
  int blocks = max_blocks > 0 ? max_blocks : 148;
  const long long per = (((n + 3) >> 2) + world - 1) / world;
  const long long need = (per + 255) / 256;
  if (need < blocks) blocks = (int)(need < 1 ? 1 : need);
  p2p_adamw_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(peers, rank, world, exp_avg, exp_avg_sq, n, (unsigned int)step, k, ticket);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}
