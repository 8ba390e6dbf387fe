// Dense layer contraction, IEEE fp32 on the CUDA cores ("fp32" precision mode).
//
//   C[M,N] = epilogue( normA(op_a(A))[M,K] . op_b(B)[K,N] )
//
// This is the exact-fp32 parity path for the three ReLU MLPs of the reference (actor
// 264-1024-1024-512-29, critic ...-1, discriminator 114-1024-512-1; configs/agent/add_g1.yaml:2-9,
// nets/fc_3layers_1024units.py, nets/fc_2layers_1024units.py) -- forward  (A=[M,K] activations,
// B=nn.Linear weight [N,K]), input gradient (B=[K,N]) and weight gradient (A transposed, split-K).
// The tensor-core modes ("tf32x3", "tf32", "bf16") live in gemm_tc.cu and share this argument struct.
//
// Tile 128x128x16, 256 threads, 8x8 accumulators per thread, register-prefetched double-buffered
// shared tiles, 128-bit global loads whenever the operand's contiguous dimension allows it.
#include "common.cuh"
#include "addk.h"

int addk_gemm_tc(cudaStream_t st, const addk_gemm_args& a, int precision);

namespace addk {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4, NT = 256;

struct GemmParams {
  const float* A; const float* B; float* C;
  int lda, ldb, ldc, M, N, K;
  const float* bias; const float* a_mean; const float* a_std; const float* mask; int ld_mask;
  int relu, accumulate, k_chunk;
  long long slab_stride;
};

// Load one BKxBM (or BKxBN) operand tile into registers.
//  KCONTIG = true : memory is [rows, K] with K contiguous (rows = M or N index)   -> 2 x float4 along K
//  KCONTIG = false: memory is [K, rows] with rows contiguous                       -> 2 x float4 along rows
template <bool KCONTIG, bool VEC, bool NORM>
__device__ __forceinline__ void load_tile(const float* __restrict__ P, int ld, int row0, int rows, int k0, int k1,
                                          const float* __restrict__ mean, const float* __restrict__ sd, float (&r)[8]) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int idx = tid + h * NT;
    if (KCONTIG) {
      const int row = row0 + idx / 4, k = k0 + (idx % 4) * 4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (row < rows) {
        const float* src = P + (size_t)row * ld + k;
        if (VEC && k + 3 < k1) {
          float4 q = *reinterpret_cast<const float4*>(src);
          v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) if (k + i < k1) v[i] = src[i];
        }
        if (NORM) {
#pragma unroll
          for (int i = 0; i < 4; ++i) if (k + i < k1) v[i] = (v[i] - mean[k + i]) / sd[k + i];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) r[h * 4 + i] = v[i];
    } else {
      const int k = k0 + idx / 32, row = row0 + (idx % 32) * 4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (k < k1) {
        const float* src = P + (size_t)k * ld + row;
        if (VEC && row + 3 < rows) {
          float4 q = *reinterpret_cast<const float4*>(src);
          v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) if (row + i < rows) v[i] = src[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) r[h * 4 + i] = v[i];
    }
  }
}

template <bool KCONTIG>
__device__ __forceinline__ void store_tile(float (*S)[BM + PAD], const float (&r)[8]) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int idx = tid + h * NT;
    if (KCONTIG) {
      const int row = idx / 4, k = (idx % 4) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) S[k + i][row] = r[h * 4 + i];
    } else {
      const int k = idx / 32, row = (idx % 32) * 4;
      *reinterpret_cast<float4*>(&S[k][row]) = make_float4(r[h * 4], r[h * 4 + 1], r[h * 4 + 2], r[h * 4 + 3]);
    }
  }
}

template <bool A_T, bool B_T, bool VEC_A, bool VEC_B, bool NORM>
__global__ void __launch_bounds__(NT, 2) sgemm_kernel(const GemmParams p) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kb = blockIdx.z * p.k_chunk;
  const int ke = min(p.K, kb + p.k_chunk);
  float* C = p.C + (size_t)blockIdx.z * p.slab_stride;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float ra[8], rb[8];
  // A operand: not transposed -> [M,K] K contiguous ; transposed -> [K,M] M contiguous
  // B operand: B_T -> [N,K] K contiguous ; else [K,N] N contiguous
  load_tile<!A_T, VEC_A, NORM>(p.A, p.lda, m0, p.M, kb, ke, p.a_mean, p.a_std, ra);
  load_tile<B_T, VEC_B, false>(p.B, p.ldb, n0, p.N, kb, ke, nullptr, nullptr, rb);
  store_tile<!A_T>(As[0], ra);
  store_tile<B_T>(Bs[0], rb);
  __syncthreads();
  int buf = 0;
  for (int k0 = kb; k0 < ke; k0 += BK) {
    const bool more = k0 + BK < ke;
    if (more) {
      load_tile<!A_T, VEC_A, NORM>(p.A, p.lda, m0, p.M, k0 + BK, ke, p.a_mean, p.a_std, ra);
      load_tile<B_T, VEC_B, false>(p.B, p.ldb, n0, p.N, k0 + BK, ke, nullptr, nullptr, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      store_tile<!A_T>(As[buf ^ 1], ra);
      store_tile<B_T>(Bs[buf ^ 1], rb);
      __syncthreads();
      buf ^= 1;
    }
  }
  // epilogue
  const bool vec_c = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) &&
                     (!p.mask || (((p.ld_mask & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.mask) & 15) == 0)));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + (jh == 0 ? tx * 4 : 64 + tx * 4);
      if (n >= p.N) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x = acc[i][jh * 4 + j];
        if (p.bias && n + j < p.N) x += p.bias[n + j];
        if (p.relu) x = fmaxf(x, 0.f);
        v[j] = x;
      }
      float* dst = C + (size_t)m * p.ldc + n;
      if (vec_c && n + 3 < p.N) {
        if (p.mask) {
          float4 mk = *reinterpret_cast<const float4*>(p.mask + (size_t)m * p.ld_mask + n);
          v[0] = mk.x > 0.f ? v[0] : 0.f; v[1] = mk.y > 0.f ? v[1] : 0.f;
          v[2] = mk.z > 0.f ? v[2] : 0.f; v[3] = mk.w > 0.f ? v[3] : 0.f;
        }
        if (p.accumulate) {
          float4 o = *reinterpret_cast<const float4*>(dst);
          v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
        }
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (n + j >= p.N) break;
          float x = v[j];
          if (p.mask) x = p.mask[(size_t)m * p.ld_mask + n + j] > 0.f ? x : 0.f;
          if (p.accumulate) x += dst[j];
          dst[j] = x;
        }
      }
    }
  }
}

template <bool A_T, bool B_T, bool VA, bool VB>
static void launch_norm(const GemmParams& p, dim3 grid, cudaStream_t st) {
  if (p.a_mean) sgemm_kernel<A_T, B_T, VA, VB, true><<<grid, NT, 0, st>>>(p);
  else sgemm_kernel<A_T, B_T, VA, VB, false><<<grid, NT, 0, st>>>(p);
}
template <bool A_T, bool B_T>
static void launch_vec(const GemmParams& p, dim3 grid, cudaStream_t st, bool va, bool vb) {
  if (va && vb) launch_norm<A_T, B_T, true, true>(p, grid, st);
  else if (va) launch_norm<A_T, B_T, true, false>(p, grid, st);
  else if (vb) launch_norm<A_T, B_T, false, true>(p, grid, st);
  else launch_norm<A_T, B_T, false, false>(p, grid, st);
}

int sgemm_launch(cudaStream_t st, const addk_gemm_args& a) {
  GemmParams p;
  p.A = a.A; p.B = a.B; p.C = a.C; p.lda = a.lda; p.ldb = a.ldb; p.ldc = a.ldc;
  p.M = a.M; p.N = a.N; p.K = a.K; p.bias = a.bias; p.a_mean = a.a_mean; p.a_std = a.a_std;
  p.mask = a.relu_mask_src; p.ld_mask = a.ld_mask; p.relu = a.relu; p.accumulate = a.accumulate;
  int split = a.split_k > 1 ? a.split_k : 1;
  int chunk = (a.K + split - 1) / split;
  chunk = ((chunk + BK - 1) / BK) * BK;
  p.k_chunk = chunk;
  p.slab_stride = a.slab_stride > 0 ? a.slab_stride : (long long)a.M * a.ldc;
  if (a.a_mean && a.trans_a) return ADDK_ERR_UNSUPPORTED;
  if (split > 1 && (a.bias || a.relu || a.relu_mask_src || a.accumulate)) return ADDK_ERR_ARG;
  dim3 grid((a.N + BN - 1) / BN, (a.M + BM - 1) / BM, split);
  const bool va = ((a.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.A) & 15) == 0);
  const bool vb = ((a.ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.B) & 15) == 0);
  if (!a.trans_a && a.trans_b) launch_vec<false, true>(p, grid, st, va, vb);
  else if (!a.trans_a && !a.trans_b) launch_vec<false, false>(p, grid, st, va, vb);
  else if (a.trans_a && !a.trans_b) launch_vec<true, false>(p, grid, st, va, vb);
  else launch_vec<true, true>(p, grid, st, va, vb);
  return ADDK_OK;
}

}  // namespace addk

extern "C" int addk_gemm(void* stream, const addk_gemm_args* a, int precision) {
  if (!a || !a->A || !a->B || !a->C || a->M <= 0 || a->N <= 0 || a->K <= 0) return ADDK_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (precision == 0 && (a->relu_bits_in || a->relu_bits_out || a->colsum_partials)) { addk_set_error("gemm: relu_bits_* / colsum_partials need precision f16x3 or bf16"); return ADDK_ERR_ARG; }
  if (precision == 0) rc = addk::sgemm_launch(st, *a);
  else rc = addk_gemm_tc(st, *a, precision);
  if (rc != ADDK_OK) return rc;
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}
