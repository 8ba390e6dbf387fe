// Dense layer contraction on the 5th-generation tensor cores: TMA-fed tcgen05.mma tiles with the
// accumulator in TMEM ("tf32" and "tf32x3" precision modes of addk_gemm; argument struct shared with gemm.cu).
//
//   C[M,N] = epilogue( op_a(A)[M,K] . op_b(B)[K,N] ),  fp32 operands in HBM, fp32 accumulate, fp32 out
//
// One CTA = one 128 x BN output tile (BN = 256 / 128 / 64), 192 threads, warp-specialised:
//   warp 0      TMA producer: cp.async.bulk.tensor.2d of the A / B k-blocks (32 fp32 = 128 B wide,
//               SWIZZLE_128B) into a ring of shared-memory stages, completion on mbarriers;
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (kind::tf32, M=128, N=BN, K=8 per
//               instruction), tcgen05.commit releases the stage / publishes the accumulator;
//   warps 2..5  "tf32x3" only: split each landed fp32 tile in place into hi = tf32(x) and lo = x - hi
//               (3 MMAs per k-step: lo.hi + hi.lo + hi.hi -> fp32-class accuracy, the reference's fp32
//               MLP parity mode); then the epilogue: tcgen05.ld the accumulator (each warp its own 32
//               TMEM lanes), bias / ReLU / ReLU-mask / accumulate, 128-bit stores.
// Operands may be K-major (row = M or N index, contraction contiguous: forward layers) or MN-major
// (row = contraction index: input-gradient and weight-gradient GEMMs); both are described to the tensor
// core through shared-memory matrix descriptors, no transposed copies are made.  Split-K (weight
// gradients, contraction over the minibatch rows) writes one partial slab per blockIdx.z.
//
// Every mbarrier wait is bounded by a clock watchdog that traps instead of hanging the GPU.
#include "common.cuh"
#include "addk.h"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <cuda_fp16.h>
#include "switches.h"
#include "h3_scale.cuh"

extern int g_addk_last_gemm_kernel;      // api.cu: id of the kernel the last addk_gemm dispatched to (tests)
extern long long* g_addk_stamps;         // api.cu: optional device buffer for clock stamps (addk_debug_set_stamp_buffer)

namespace addk { int sgemm_launch(cudaStream_t st, const addk_gemm_args& a); }

namespace addk_tc {

static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

constexpr int BM = 128;        // UMMA M
constexpr int NTHREADS = 192;

struct Params {
  float* C; int ldc; int M, N, K;
  const float* bias; const float* mask; int ld_mask;
  int relu, accumulate;
  int kb_per_split;            // k-blocks per blockIdx.z
  int a_mn, b_mn;              // operand is MN-major (memory rows = contraction index)
  long long slab_stride;       // floats between split-K slabs
  void* C16;                   // bf16 kernel: optional bf16 copy of the output (same leading dimension)
  long long* dbg;              // experiment: clock64 stamps of CTA 0 (addk_debug_set_stamp_buffer; NULL = off)
  int pair_flags;              // CTA-pair kernel experiments: bit0 cluster-scope waits, bit1 relaxed remote arrives
  uint32_t* c_amax;            // f16x3 kernels: slot {W, max} of C: atomicMax of the bit patterns of |C| as stored into [1]
  uint16_t* c_hi;              // f16x3 persistent kernel: fp16 planes of C written by the epilogue with the scale of the
  long long c_plane;           //   slot's sticky word W (NULL: not wanted); lo plane c_plane elements after the hi plane
  float c_scale;               // device-side only: filled in by the kernel
  int c16_in_staged;           // persistent kernel in bf16 mode: the shared epilogue helpers also write the bf16 copy C16
  int no_f32;                  // persistent kernel: the fp32 output is NOT written (C lives only as its 16-bit copy / planes)
  const uint16_t* mask16;      // persistent kernel: ReLU-mask source as 16-bit values (bf16 copy or fp16 hi plane, pitch ld_mask):
                               //   element > 0  <=>  sign bit clear and magnitude bits non-zero; used instead of `mask`
  // ReLU masks as bit planes (persistent f16x3 kernel, N % 128 == 0): bit c % 32 of word [row * ld_bits + c / 32]
  uint32_t* bits_out = nullptr;        // written by the store loop of a layer: output element > 0
  const uint32_t* bits_in = nullptr;   // applied in the accumulator's own layout (lane = row) before staging: replaces mask / mask16
  int ld_bits = 0;                     // words per row (a multiple of 4)
  float* colpart = nullptr;            // persistent kernels: column sums per 32-row block of the output, [ceil(M / 32), N]
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: ~2 s at 2 GHz, then trap (a launch error instead of a hung device)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// Shared-memory matrix descriptor (descriptor version 1 = Blackwell).  layout: 2 = SWIZZLE_128B (K-major tiles),
// 1 = SWIZZLE_128B with 32-byte atoms -- the only swizzle the tensor core accepts for MN-major 32-bit operands.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// one lane of a converged warp (always the same one for a full mask)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// f16x3: x * s -> (hi, lo) fp16 pair, four at a time (see the f16x3 section below)
__device__ __forceinline__ void h3_split4(const float4& o, float s, uint2& h, uint2& l) {
  const float x0 = o.x * s, x1 = o.y * s, x2 = o.z * s, x3 = o.w * s;
  const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
  const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
  const __half2 l01 = __floats2half2_rn(x0 - f01.x, x1 - f01.y), l23 = __floats2half2_rn(x2 - f23.x, x3 - f23.y);
  h = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
  l = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}
// The epilogue fields a worker warp decides for itself (the rest of Params is read from the kernel parameters where it is
// used: a whole worker-local copy of Params stayed live across the drain loop and pushed it over the register budget)
struct EpiOv { uint16_t* c_hi; float c_scale; const float* bias; int relu; };
__device__ __forceinline__ void h3_emit1(const Params& p, const EpiOv& e, size_t off, float v) {
  const float xs = v * e.c_scale;
  const __half h = __float2half_rn(xs);
  e.c_hi[off] = __half_as_ushort(h);
  e.c_hi[p.c_plane + off] = __half_as_ushort(__float2half_rn(xs - __half2float(h)));
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// Epilogue helpers shared by the fp32-output kernels.  A warp owns 32 accumulator rows (one per lane, the TMEM view) x CW
// columns.  It first writes them into a swizzled staging tile in shared memory (stage_put: slot = float4 index inside
// the row, XOR-ed with row & 7 -> conflict-free both ways), then store_staged() walks the tile ROW by row so that one
// store instruction covers min(CW, 128) consecutive floats of one output row (512 contiguous bytes), with bias / ReLU /
// ReLU-mask / accumulate applied on that coalesced side.  Measured with clock stamps on the 16384x1024x1024 layer: the
// earlier shape (4 rows x 128 B per instruction, every warp of every CTA on the same 128-byte column phase at the same
// time) kept address bits 7-8 constant GPU-wide and drained at 2.7 B/clk/SM (38k of a CTA's 108k cycles); full rows
// take 11k.
// ReLU mask of four consecutive elements from 16-bit values (bf16 or fp16: same sign / zero encoding): 1.0 where > 0
__device__ __forceinline__ float4 mask4_from16(const uint16_t* m) {
  const uint2 r = __ldg(reinterpret_cast<const uint2*>(m));
  const uint32_t a = r.x & 0xFFFFu, b = r.x >> 16, c = r.y & 0xFFFFu, d = r.y >> 16;
  return make_float4((a - 1u) < 0x7FFFu ? 1.f : 0.f, (b - 1u) < 0x7FFFu ? 1.f : 0.f, (c - 1u) < 0x7FFFu ? 1.f : 0.f,
                     (d - 1u) < 0x7FFFu ? 1.f : 0.f);          // 0x0001 .. 0x7FFF: positive, non-zero (NaN cannot occur here)
}
__device__ __forceinline__ bool mask1_from16(const uint16_t* m) { return ((uint32_t)__ldg(m) - 1u) < 0x7FFFu; }

// ReLU bit planes (persistent f16x3 kernel).  Both sides work in the accumulator's own layout, where a lane owns one row
// and the warp's 128 columns: the ReLU layer adds the bias, clamps and collects "> 0" bits BEFORE staging (four words per
// lane and tile, one 16-byte store), the masked layer loads the same four words and zeroes its accumulators before
// staging.  Layout of a row's words: column 64 g + 4 i + k (i < 16, k < 4) -> word 2 g + (k >> 1), bit 16 (k & 1) + i.
// (First versions produced the bits in the row-major store loop -- a shuffle reduction, then four ballots per store
// instruction: +28 instructions per 16-byte store and 4.5k cycles per tile on every ReLU layer.)
template <int CW>
__device__ __forceinline__ void stage_put(float4* stg, int lane, int slot, float a, float b, float c, float d) {
  stg[lane * (CW / 4) + (slot ^ (lane & 7))] = make_float4(a, b, c, d);
}

template <int CW>
__device__ __forceinline__ void store_staged(const Params& p, const EpiOv& e, float* Cz, const float4* stg, int lane, int grow0, int col0,
                                             float* vmax) {
  constexpr int S = CW / 4;                        // float4 slots per row
  constexpr int RPI = S >= 32 ? 1 : 32 / S;        // rows per store instruction
  constexpr int PPR = S > 32 ? S / 32 : 1;         // instructions per row
  constexpr int NIT = (32 / RPI) * PPR;
  static_assert(S % 8 == 0 && NIT % 8 == 0, "tile shape");
  const int sub_r = S >= 32 ? 0 : lane / S;
  const int sl0 = S >= 32 ? lane : lane % S;
  float4 b4[PPR];
#pragma unroll
  for (int ps = 0; ps < PPR; ++ps) {
    const int col = col0 + 4 * (sl0 + 32 * ps);
    b4[ps] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e.bias) {
      if (col + 3 < p.N) b4[ps] = *reinterpret_cast<const float4*>(e.bias + col);
      else { if (col < p.N) b4[ps].x = e.bias[col]; if (col + 1 < p.N) b4[ps].y = e.bias[col + 1]; if (col + 2 < p.N) b4[ps].z = e.bias[col + 2]; }
    }
  }
#pragma unroll 1
  for (int i0 = 0; i0 < NIT; i0 += 8) {
    float4 m4[8], a4[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {                  // operand loads of eight instructions first
      const int it = i0 + u;
      const int grow = grow0 + (it / PPR) * RPI + sub_r;
      const int col = col0 + 4 * (sl0 + 32 * (it % PPR));
      m4[u] = make_float4(1.f, 1.f, 1.f, 1.f);
      a4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (grow < p.M && col + 3 < p.N) {
        if (p.mask16) m4[u] = mask4_from16(p.mask16 + (size_t)grow * p.ld_mask + col);
        else if (p.mask) m4[u] = *reinterpret_cast<const float4*>(p.mask + (size_t)grow * p.ld_mask + col);
        if (p.accumulate) a4[u] = *reinterpret_cast<const float4*>(Cz + (size_t)grow * p.ldc + col);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int it = i0 + u;
      const int r = (it / PPR) * RPI + sub_r;
      const int sl = sl0 + 32 * (it % PPR);
      const int grow = grow0 + r, col = col0 + 4 * sl;
      float4 o = stg[r * S + (sl ^ (r & 7))];
      if (grow < p.M && col < p.N) {
        const float4 bb = b4[it % PPR];
        o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
        if (e.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        float* dstp = Cz + (size_t)grow * p.ldc + col;
        if (col + 3 < p.N) {
          o.x = m4[u].x > 0.f ? o.x : 0.f; o.y = m4[u].y > 0.f ? o.y : 0.f;
          o.z = m4[u].z > 0.f ? o.z : 0.f; o.w = m4[u].w > 0.f ? o.w : 0.f;
          o.x += a4[u].x; o.y += a4[u].y; o.z += a4[u].z; o.w += a4[u].w;
          if (!p.no_f32) *reinterpret_cast<float4*>(dstp) = o;
          if (vmax) *vmax = fmaxf(fmaxf(*vmax, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
          if (e.c_hi) {
            uint2 h, l;
            h3_split4(o, e.c_scale, h, l);
            const size_t off = (size_t)grow * p.ldc + col;
            *reinterpret_cast<uint2*>(e.c_hi + off) = h;
            *reinterpret_cast<uint2*>(e.c_hi + p.c_plane + off) = l;
          }
          if (p.C16 && p.c16_in_staged)
            *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.C16) + (size_t)grow * p.ldc + col) =
                make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        } else {
          const float oo[4] = {o.x, o.y, o.z, o.w};
          for (int q4 = 0; q4 < 4 && col + q4 < p.N; ++q4) {
            float xv = oo[q4];
            if (p.mask16) xv = mask1_from16(p.mask16 + (size_t)grow * p.ld_mask + col + q4) ? xv : 0.f;
            else if (p.mask) xv = p.mask[(size_t)grow * p.ld_mask + col + q4] > 0.f ? xv : 0.f;
            if (p.accumulate) xv += dstp[q4];
            if (!p.no_f32) dstp[q4] = xv;
            if (vmax) *vmax = fmaxf(*vmax, fabsf(xv));
            if (e.c_hi) h3_emit1(p, e, (size_t)grow * p.ldc + col + q4, xv);
            if (p.C16 && p.c16_in_staged) reinterpret_cast<uint16_t*>(p.C16)[(size_t)grow * p.ldc + col + q4] = (uint16_t)(pack_bf16x2(xv, 0.f) & 0xFFFFu);
          }
        }
      }
    }
  }
}

// Column sums of a 32-row block (p.colpart): the row-major walk leaves, in every lane, the sum of its four columns over
// the rows it visited (lanes l and l + S hold the other half of the rows): add the halves, one float4 store per column quad.
template <int S>
__device__ __forceinline__ void colpart_put(const Params& p, float4 cs, int lane, int grow0, int col) {
#pragma unroll
  for (int d = S; d < 32; d <<= 1) {
    cs.x += __shfl_xor_sync(0xffffffffu, cs.x, d); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, d);
    cs.z += __shfl_xor_sync(0xffffffffu, cs.z, d); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, d);
  }
  if (lane < S) *reinterpret_cast<float4*>(p.colpart + (size_t)(grow0 >> 5) * p.N + col) = cs;
}

// The same walk for a sub-tile that lies completely inside the matrix and does not accumulate: no bounds checks, one
// pointer increment per store.  (ncu on the persistent f16x3 kernel: the general version above executes ~300
// instructions per 16-byte store -- 64-bit address arithmetic, constant-bank reloads and tail predicates -- which made
// the epilogue of a 128 x 256 tile 17.7k cycles and instruction-bound.)
template <int CW, bool MASK, bool PLANES, bool BF16OUT = false>
__device__ __forceinline__ void store_staged_interior(const Params& p, const EpiOv& e, float* __restrict__ Cz, const float4* stg, int lane,
                                                      int grow0, int col0, float* vmax) {
  constexpr int S = CW / 4;                        // float4 slots per row
  static_assert(S <= 32 && 32 % S == 0, "one store instruction covers 32 / S whole rows");
  constexpr int RPI = 32 / S;
  const int sub_r = lane / S, sl = lane % S;
  const int col = col0 + 4 * sl;
  float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (e.bias) bb = __ldg(reinterpret_cast<const float4*>(e.bias + col));
  const bool relu = e.relu != 0;
  float* dst = Cz + (size_t)(grow0 + sub_r) * p.ldc + col;
  uint16_t* dhi = PLANES ? e.c_hi + (size_t)(grow0 + sub_r) * p.ldc + col : nullptr;
  uint16_t* d16 = BF16OUT ? reinterpret_cast<uint16_t*>(p.C16) + (size_t)(grow0 + sub_r) * p.ldc + col : nullptr;
  const float cs = e.c_scale;
  const bool m16 = MASK && p.mask16 != nullptr;       // uniform
  const float* mk = (MASK && !m16) ? p.mask + (size_t)(grow0 + sub_r) * p.ld_mask + col : nullptr;
  const uint16_t* mk16 = m16 ? p.mask16 + (size_t)(grow0 + sub_r) * p.ld_mask + col : nullptr;
  const size_t dstep = (size_t)RPI * p.ldc, mstep = MASK ? (size_t)RPI * p.ld_mask : 0;
  const bool f32 = !p.no_f32;
  float vm = 0.f;
  float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);      // column sums of this lane's four columns over its 16 rows (p.colpart)
  constexpr int EB = MASK ? 8 : 4;       // store instructions per loop iteration (masked: eight mask loads in flight)
  // (a real loop: fully unrolled, the store walk of one pass was 13 KB of straight-line code that every warp streamed
  //  through the instruction caches once per tile -- ncu: "no instruction" was the top stall reason of the epilogue)
#pragma unroll 1
  for (int i0 = 0; i0 < 32 / RPI; i0 += EB) {
    float4 m4[EB];
    if (MASK) {
      if (m16) {
#pragma unroll
        for (int u = 0; u < EB; ++u) m4[u] = mask4_from16(mk16 + (size_t)(i0 + u) * mstep);
      } else {
#pragma unroll
        for (int u = 0; u < EB; ++u) m4[u] = __ldg(reinterpret_cast<const float4*>(mk + (size_t)(i0 + u) * mstep));
      }
    }
#pragma unroll
    for (int u = 0; u < EB; ++u) {
      const int r = (i0 + u) * RPI + sub_r;
      float4 o = stg[r * S + (sl ^ (r & 7))];
      o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
      if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      if (MASK) {
        o.x = m4[u].x > 0.f ? o.x : 0.f; o.y = m4[u].y > 0.f ? o.y : 0.f;
        o.z = m4[u].z > 0.f ? o.z : 0.f; o.w = m4[u].w > 0.f ? o.w : 0.f;
      }
      if (f32) *reinterpret_cast<float4*>(dst) = o;
      dst += dstep;
      csum.x += o.x; csum.y += o.y; csum.z += o.z; csum.w += o.w;
      if (PLANES) {
        uint2 h, l;
        h3_split4(o, cs, h, l);
        *reinterpret_cast<uint2*>(dhi) = h;
        *reinterpret_cast<uint2*>(dhi + p.c_plane) = l;
        dhi += dstep;
      }
      if (BF16OUT) {
        *reinterpret_cast<uint2*>(d16) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        d16 += dstep;
      }
      vm = fmaxf(fmaxf(vm, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
    }
  }
  if (vmax) *vmax = fmaxf(*vmax, vm);
  if (p.colpart) colpart_put<S>(p, csum, lane, grow0, col);
}

// A sub-tile whose columns lie inside the matrix but whose LAST rows do not (M = 16385: the discriminator chain's extra
// row makes a 65th row tile with one valid row): one compact, non-unrolled loop over the valid row pairs with every
// epilogue option decided at run time.  (The general walk above spends ~10k instructions per warp on such a sub-tile
// whatever the number of valid rows; as the tail of a persistent kernel that was ~20 us per layer.)
template <int CW>
__device__ __forceinline__ void store_staged_rows(const Params& p, const EpiOv& e, float* __restrict__ Cz, const float4* stg, int lane, int grow0,
                                                  int col0, int nrows, float* vmax) {
  constexpr int S = CW / 4, RPI = 32 / S;
  const int sub_r = lane / S, sl = lane % S, col = col0 + 4 * sl;
  float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (e.bias) bb = __ldg(reinterpret_cast<const float4*>(e.bias + col));
  float vm = 0.f;
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
  for (int i = 0; i * RPI < nrows; ++i) {          // warp-uniform bound
    const int r = i * RPI + sub_r;
    const bool ok = r < nrows;
    const size_t grow = (size_t)(grow0 + r);
    float4 o = stg[r * S + (sl ^ (r & 7))];
    o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
    if (e.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    if (ok) {
      if (p.mask16 || p.mask) {
        const float4 m4 = p.mask16 ? mask4_from16(p.mask16 + grow * p.ld_mask + col)
                                   : __ldg(reinterpret_cast<const float4*>(p.mask + grow * p.ld_mask + col));
        o.x = m4.x > 0.f ? o.x : 0.f; o.y = m4.y > 0.f ? o.y : 0.f; o.z = m4.z > 0.f ? o.z : 0.f; o.w = m4.w > 0.f ? o.w : 0.f;
      }
      const size_t off = grow * p.ldc + col;
      cs.x += o.x; cs.y += o.y; cs.z += o.z; cs.w += o.w;
      if (!p.no_f32) *reinterpret_cast<float4*>(Cz + off) = o;
      if (e.c_hi) {
        uint2 h, l;
        h3_split4(o, e.c_scale, h, l);
        *reinterpret_cast<uint2*>(e.c_hi + off) = h;
        *reinterpret_cast<uint2*>(e.c_hi + p.c_plane + off) = l;
      }
      if (p.C16 && p.c16_in_staged)
        *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.C16) + off) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
      vm = fmaxf(fmaxf(vm, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
    }
  }
  if (vmax) *vmax = fmaxf(*vmax, vm);
  if (p.colpart) colpart_put<S>(p, cs, lane, grow0, col);
}

// row-per-lane scalar fallback for outputs that are not 16-byte aligned (ldc % 4 != 0)
__device__ __forceinline__ void store_row_scalar(const Params& p, const EpiOv& e, float* Cz, int row, int cbase, const float* acc32,
                                                 float* vmax) {
  float* dstp = Cz + (size_t)row * p.ldc + cbase;
  const float* mk = p.mask ? p.mask + (size_t)row * p.ld_mask + cbase : nullptr;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int col = cbase + j;
    if (col < p.N) {
      float xv = acc32[j];
      if (e.bias) xv += e.bias[col];
      if (e.relu) xv = fmaxf(xv, 0.f);
      if (p.mask16) xv = mask1_from16(p.mask16 + (size_t)row * p.ld_mask + col) ? xv : 0.f;
      else if (mk) xv = mk[j] > 0.f ? xv : 0.f;
      if (p.accumulate) xv += dstp[j];
      if (!p.no_f32) dstp[j] = xv;
      if (vmax) *vmax = fmaxf(*vmax, fabsf(xv));
      if (e.c_hi) h3_emit1(p, e, (size_t)row * p.ldc + col, xv);
      if (p.C16 && p.c16_in_staged) reinterpret_cast<uint16_t*>(p.C16)[(size_t)row * p.ldc + col] = (uint16_t)(pack_bf16x2(xv, 0.f) & 0xFFFFu);
    }
  }
}
// the kernels that take every epilogue field from Params
template <int CW>
__device__ __forceinline__ void store_staged(const Params& p, float* Cz, const float4* stg, int lane, int grow0, int col0,
                                             float* vmax = nullptr) {
  store_staged<CW>(p, EpiOv{p.c_hi, p.c_scale, p.bias, p.relu}, Cz, stg, lane, grow0, col0, vmax);
}
__device__ __forceinline__ void store_row_scalar(const Params& p, float* Cz, int row, int cbase, const float* acc32, float* vmax = nullptr) {
  store_row_scalar(p, EpiOv{p.c_hi, p.c_scale, p.bias, p.relu}, Cz, row, cbase, acc32, vmax);
}
__device__ __forceinline__ bool epilogue_vec_ok(const Params& p, const float* Cz) {
  return ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(Cz) & 15) == 0) &&
         (!p.bias || ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0)) &&
         (!p.mask || (((p.ld_mask & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.mask) & 15) == 0))) &&
         (!p.mask16 || (((p.ld_mask & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.mask16) & 7) == 0)));
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static bool resolve_encode() {
  if (g_encode) return true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return false;
  g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  return true;
}

// tf32x3 drains share these with the f16x3 kernels
constexpr int X3_THREADS = 320;

// ---------------------------------------------------------------------------------------------------------------
// bf16 kernel (precision "bf16", BASELINE config 4): bf16 operands in HBM (the producers write a bf16 twin of every
// activation / gradient / weight), tcgen05.mma.kind::f16 (M=128, N=BN, K=16), fp32 accumulation in TMEM, fp32 and/or
// bf16 output.  64-element (128-byte) k-blocks; K-major tiles use SWIZZLE_128B, MN-major tiles the standard 16-bit
// MN-major SWIZZLE_128B atoms (64 MN x 8 k), one TMA box = 64 MN x 64 k = 8 KB.  6 warps: TMA producer, MMA issuer,
// 4 epilogue warps.
// ---------------------------------------------------------------------------------------------------------------
template <int BN>
struct CfgH {
  static constexpr int BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 6 ? 6 : (200 * 1024) / STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr uint32_t MN_BOX_BYTES = BK * 128u;     // 64 k-rows x 128 bytes (64 bf16 along MN)
};

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
  using C = CfgH<BN>;
  constexpr int BK = C::BK, UK = 16;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* const base_ptr = smem_raw + (base - raw_u32);
  const uint32_t bars = base + C::STAGES * C::STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
  const uint32_t tmem_full_bar = bars + 8u * (2 * C::STAGES);
  const uint32_t tmem_ptr_addr = bars + 8u * (2 * C::STAGES + 1);
  auto a_t = [&](int s) { return base + (uint32_t)s * C::STAGE_BYTES; };
  auto b_t = [&](int s) { return a_t(s) + C::A_BYTES; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kb_total = (p.K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(kb_total, kb_begin + p.kb_per_split);
  const int num_kb = kb_end - kb_begin;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"((uint32_t)C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr) : "memory");

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % C::STAGES;
        const uint32_t ph = (uint32_t)(i / C::STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), C::A_BYTES + C::B_BYTES);
        const int k0 = (kb_begin + i) * BK;
        if (!p.a_mn) {
          tma_load_2d(a_t(s), &tmA, full_bar(s), k0, m0);                       // box {64 k, 128 rows}
        } else {
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) tma_load_2d(a_t(s) + j * C::MN_BOX_BYTES, &tmA, full_bar(s), m0 + 64 * j, k0);  // box {64 m, 64 k}
        }
        if (!p.b_mn) {
          tma_load_2d(b_t(s), &tmB, full_bar(s), k0, n0);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(b_t(s) + j * C::MN_BOX_BYTES, &tmB, full_bar(s), n0 + 64 * j, k0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D fp32, A/B bf16
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.a_mn ? 1 : 0) << 15) |
                             ((uint32_t)(p.b_mn ? 1 : 0) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      // K-major: 128-byte rows, 8-row groups 1024 B apart, 16 k = 32 B.  MN-major: 64-MN atoms one box (8 KB) apart (LBO),
      // 8-k groups 1024 B apart (SBO), 16 k = 2048 B.
      const uint32_t a_lbo = p.a_mn ? C::MN_BOX_BYTES : 16u, b_lbo = p.b_mn ? C::MN_BOX_BYTES : 16u;
      const uint32_t a_kstep = p.a_mn ? 2048u : 32u, b_kstep = p.b_mn ? 2048u : 32u;
      uint32_t acc = 0;
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % C::STAGES;
        const uint32_t ph = (uint32_t)(i / C::STAGES) & 1u;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < BK / UK; ++ks) {
          umma_bf16(tmem_base, smem_desc(a_t(s) + ks * a_kstep, a_lbo, 1024u, 2u), smem_desc(b_t(s) + ks * b_kstep, b_lbo, 1024u, 2u),
                    idesc, acc);
          acc = 1;
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp & 3;
    float* Cz = p.C ? p.C + (size_t)blockIdx.z * p.slab_stride : nullptr;
    uint16_t* C16 = reinterpret_cast<uint16_t*>(p.C16);
    const bool vec = ((p.ldc & 3) == 0) && (!Cz || (reinterpret_cast<uintptr_t>(Cz) & 15) == 0) &&
                     (!C16 || (reinterpret_cast<uintptr_t>(C16) & 7) == 0) &&
                     (!p.bias || ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0)) &&
                     (!p.mask || (((p.ld_mask & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.mask) & 15) == 0)));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    float4* stg = reinterpret_cast<float4*>(base_ptr + 4096 * q);
    const int l_row = lane >> 3, l_c4 = lane & 7;
    const int row = m0 + 32 * q + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= p.N) break;
      float4 m4[8];
      const int colv = n0 + c0 + 4 * l_c4;
      const bool full4 = vec && (colv + 3 < p.N);
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int grow = m0 + 32 * q + it * 4 + l_row;
        m4[it] = make_float4(1.f, 1.f, 1.f, 1.f);
        if (p.mask && full4 && grow < p.M) m4[it] = *reinterpret_cast<const float4*>(p.mask + (size_t)grow * p.ld_mask + colv);
      }
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)c0, v);
      if (vec) {
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4)
          stg[lane * 8 + (c4 ^ (lane & 7))] = make_float4(__uint_as_float(v[4 * c4]), __uint_as_float(v[4 * c4 + 1]),
                                                          __uint_as_float(v[4 * c4 + 2]), __uint_as_float(v[4 * c4 + 3]));
        __syncwarp();
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) {
          if (full4) b4 = *reinterpret_cast<const float4*>(p.bias + colv);
          else { if (colv < p.N) b4.x = p.bias[colv]; if (colv + 1 < p.N) b4.y = p.bias[colv + 1]; if (colv + 2 < p.N) b4.z = p.bias[colv + 2]; }
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + l_row;
          const int grow = m0 + 32 * q + r;
          float4 o = stg[r * 8 + (l_c4 ^ (r & 7))];
          if (grow < p.M && colv < p.N) {
            o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
            if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (full4) {
              o.x = m4[it].x > 0.f ? o.x : 0.f; o.y = m4[it].y > 0.f ? o.y : 0.f;
              o.z = m4[it].z > 0.f ? o.z : 0.f; o.w = m4[it].w > 0.f ? o.w : 0.f;
              if (Cz) *reinterpret_cast<float4*>(Cz + (size_t)grow * p.ldc + colv) = o;
              if (C16) *reinterpret_cast<uint2*>(C16 + (size_t)grow * p.ldc + colv) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
            } else {
              const float oo[4] = {o.x, o.y, o.z, o.w};
              for (int e = 0; e < 4 && colv + e < p.N; ++e) {
                float xv = oo[e];
                if (p.mask) xv = p.mask[(size_t)grow * p.ld_mask + colv + e] > 0.f ? xv : 0.f;
                if (Cz) Cz[(size_t)grow * p.ldc + colv + e] = xv;
                if (C16) C16[(size_t)grow * p.ldc + colv + e] = (uint16_t)(pack_bf16x2(xv, 0.f) & 0xFFFFu);
              }
            }
          }
        }
        __syncwarp();
      } else if (row < p.M) {
        for (int j = 0; j < 32; ++j) {
          const int col = n0 + c0 + j;
          if (col >= p.N) break;
          float xv = __uint_as_float(v[j]);
          if (p.bias) xv += p.bias[col];
          if (p.relu) xv = fmaxf(xv, 0.f);
          if (p.mask) xv = p.mask[(size_t)row * p.ld_mask + col] > 0.f ? xv : 0.f;
          if (Cz) Cz[(size_t)row * p.ldc + col] = xv;
          if (C16) C16[(size_t)row * p.ldc + col] = (uint16_t)(pack_bf16x2(xv, 0.f) & 0xFFFFu);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
  }
}

// 2-D bf16 tensor map: memory [outer, inner] with `ld` elements between rows; box {64, box_rows}, 128-byte swizzle.
static bool make_map_bf16(CUtensorMap* map, const void* ptr, long long inner, long long outer, long long ld, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BN>
static int launch_bf16(cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, dim3 grid) {
  using C = CfgH<BN>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_tc_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess) {
      addk_set_error("gemm_tc: cannot raise the dynamic shared memory limit");
      return ADDK_ERR_LAUNCH;
    }
    configured = true;
  }
  gemm_tc_bf16_kernel<BN><<<grid, NTHREADS, C::SMEM_BYTES, st>>>(ta, tb, p);
  return ADDK_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// f16x3 kernel (precision "f16x3"): the fp32-parity mode on the fp16 tensor pipe.
//
// Why: in tf32x3 the tensor core reads 4-byte operands -- per 16 k of a 128x256 tile the six kind::tf32 MMAs read
// 72 KB of shared memory and the hi/lo split moves another 48 KB, 120 KB against the SM's 128 B/clk: the main loop is
// shared-memory-bandwidth bound at ~200 cycles per MMA (nominal 134).  Here every operand is split ONCE, outside the
// GEMM, into two fp16 planes
//     hi = fp16_rn(x * s),   lo = fp16_rn(x * s - hi),   s = 2^(14 - floor(log2(max|x|)))
// (22+ mantissa bits for every element within 2^-17 of max|x|, an absolute error below 2^-39 max|x| for the rest; the
// residual is exact in fp32 before its rounding), and the GEMM is pure TMA -> tcgen05.mma.kind::f16 (M=128, N=BN,
// K=16, twice the tf32 rate, half the shared-memory bytes per k):
//     main  += hi_a . hi_b            (TMEM columns [0, BN))
//     cross += lo_a . hi_b + hi_a . lo_b   (TMEM columns [BN, 2 BN))
//     C = (main + cross) / (s_a * s_b)            lo.lo (2^-22 relative) is dropped
// The main accumulator is drained into fp32 registers every H3_CHUNK_KB k-blocks like in the tf32x3 kernel (the
// tensor core truncates its accumulator after every instruction).  10 warps: 0 = TMA producer, 1 = MMA issuer (whole
// warp runs the loop, one elected lane issues), 2..9 = chunk drains + epilogue.
// ---------------------------------------------------------------------------------------------------------------
template <int BN>
struct CfgH3 {
  static constexpr int BK = 32;                                // fp16 elements per k-block: 64-byte rows
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);  // [A_hi | B_hi | A_lo | B_lo]
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 6 ? 6 : (200 * 1024) / STAGE_BYTES;
  static constexpr int EPI_BYTES = 8 * 32 * (BN / 2) * 4;      // epilogue staging: 8 warps x 32 rows x BN/2 floats
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES > EPI_BYTES ? STAGES * STAGE_BYTES : EPI_BYTES;
  static constexpr int SMEM_BYTES = RING_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr uint32_t K_SBO = 8u * BK * 2u;              // K-major: 8-row groups 512 B apart, SWIZZLE_64B
  static constexpr uint32_t MN_BOX_BYTES = BK * 128u;          // MN-major: one TMA box = 64 MN (128 B) x BK k
};
constexpr int H3_CHUNK_KB = 8;                                 // drain every 256 k = 16 main-accumulator instructions

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

struct ParamsH3 {
  Params p;
  const uint32_t* a_amax; const uint32_t* b_amax;
  float comp_per_mma;          // expected truncation loss of the accumulator per accumulated instruction
};

template <int BN>
__global__ void __launch_bounds__(X3_THREADS, 1)
gemm_tc_h3_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                  const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, const ParamsH3 ph) {
  using C = CfgH3<BN>;
  const Params& p = ph.p;
  constexpr int BK = C::BK, UK = 16;
  constexpr int CPW = BN / 2, NCH = CPW / 32;
  static_assert(CPW % 32 == 0, "BN must be a multiple of 64");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* const base_ptr = smem_raw + (base - raw_u32);
  const uint32_t bars = base + C::RING_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
  const uint32_t tmem_full_bar = bars + 8u * (2 * C::STAGES);
  const uint32_t chunk_full_bar = bars + 8u * (2 * C::STAGES + 1);
  const uint32_t chunk_empty_bar = bars + 8u * (2 * C::STAGES + 2);
  const uint32_t tmem_ptr_addr = bars + 8u * (2 * C::STAGES + 3);
  auto a_hi = [&](int s) { return base + (uint32_t)s * C::STAGE_BYTES; };
  auto b_hi = [&](int s) { return a_hi(s) + C::A_BYTES; };
  auto a_lo = [&](int s) { return b_hi(s) + C::B_BYTES; };
  auto b_lo = [&](int s) { return a_lo(s) + C::A_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kb_total = (p.K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(kb_total, kb_begin + p.kb_per_split);
  const int num_kb = kb_end - kb_begin;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmAh); tma_prefetch_desc(&tmAl); tma_prefetch_desc(&tmBh); tma_prefetch_desc(&tmBl);
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    mbar_init(chunk_full_bar, 1);
    mbar_init(chunk_empty_bar, 8);                // one arrival per worker warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"((uint32_t)C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr) : "memory");

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % C::STAGES;
        const uint32_t phs = (uint32_t)(i / C::STAGES) & 1u;
        mbar_wait(empty_bar(s), phs ^ 1u);
        mbar_expect_tx(full_bar(s), C::STAGE_BYTES);
        const int k0 = (kb_begin + i) * BK;
        if (!p.a_mn) {
          tma_load_2d(a_hi(s), &tmAh, full_bar(s), k0, m0);                      // box {32 k, 128 rows}
          tma_load_2d(a_lo(s), &tmAl, full_bar(s), k0, m0);
        } else {
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) {                                    // box {64 m, 32 k}
            tma_load_2d(a_hi(s) + j * C::MN_BOX_BYTES, &tmAh, full_bar(s), m0 + 64 * j, k0);
            tma_load_2d(a_lo(s) + j * C::MN_BOX_BYTES, &tmAl, full_bar(s), m0 + 64 * j, k0);
          }
        }
        if (!p.b_mn) {
          tma_load_2d(b_hi(s), &tmBh, full_bar(s), k0, n0);
          tma_load_2d(b_lo(s), &tmBl, full_bar(s), k0, n0);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) {
            tma_load_2d(b_hi(s) + j * C::MN_BOX_BYTES, &tmBh, full_bar(s), n0 + 64 * j, k0);
            tma_load_2d(b_lo(s) + j * C::MN_BOX_BYTES, &tmBl, full_bar(s), n0 + 64 * j, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    // instruction descriptor: D fp32, A/B fp16 (format 0), majors, N>>3, M>>4
    const uint32_t idesc = (1u << 4) | ((uint32_t)(p.a_mn ? 1 : 0) << 15) | ((uint32_t)(p.b_mn ? 1 : 0) << 16) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    // K-major: 64-byte rows, SWIZZLE_64B (layout 4), 8-row groups 512 B apart, 16 k = 32 B.
    // MN-major: SWIZZLE_128B (layout 2), 64-MN atoms one box apart (LBO), 8-k groups 1024 B apart (SBO), 16 k = 2048 B.
    const uint32_t a_lbo = p.a_mn ? C::MN_BOX_BYTES : 16u, b_lbo = p.b_mn ? C::MN_BOX_BYTES : 16u;
    const uint32_t a_sbo = p.a_mn ? 1024u : C::K_SBO, b_sbo = p.b_mn ? 1024u : C::K_SBO;
    const uint32_t a_lay = p.a_mn ? 2u : 4u, b_lay = p.b_mn ? 2u : 4u;
    const uint64_t dA0 = smem_desc(a_hi(0), a_lbo, a_sbo, a_lay);
    const uint64_t dB0 = smem_desc(b_hi(0), b_lbo, b_sbo, b_lay);
    const uint64_t a_k16 = p.a_mn ? (2048u >> 4) : (32u >> 4), b_k16 = p.b_mn ? (2048u >> 4) : (32u >> 4);
    constexpr uint64_t LO16 = (C::A_BYTES + C::B_BYTES) >> 4, STAGE16 = C::STAGE_BYTES >> 4;
    const bool issuer = elect_one();
    uint32_t acc = 0, acc_x = 0, phs = 0, chunk_par = 0;
    int chunk_left = H3_CHUNK_KB;
    for (int i = 0; i < num_kb; phs ^= 1u) {
#pragma unroll
      for (int s = 0; s < C::STAGES; ++s) {
        if (i >= num_kb) break;
        mbar_wait(full_bar(s), phs);
        tc_fence_after();
        const uint64_t dah = dA0 + s * STAGE16, dbh = dB0 + s * STAGE16;
        if (issuer) {
#pragma unroll
          for (int ks = 0; ks < BK / UK; ++ks) {     // cross terms first: they overlap the drain at a chunk boundary
            umma_f16(tmem_base + BN, dah + LO16 + ks * a_k16, dbh + ks * b_k16, idesc, acc_x);
            acc_x = 1;
            umma_f16(tmem_base + BN, dah + ks * a_k16, dbh + LO16 + ks * b_k16, idesc, acc_x);
          }
        }
        if (chunk_left == 0) {                       // the workers have copied the previous chunk out of the main accumulator
          mbar_wait(chunk_empty_bar, chunk_par);
          tc_fence_after();
          acc = 0;
          chunk_par ^= 1u;
          chunk_left = H3_CHUNK_KB;
        }
        --chunk_left;
        ++i;
        if (issuer) {
#pragma unroll
          for (int ks = 0; ks < BK / UK; ++ks) {
            umma_f16(tmem_base, dah + ks * a_k16, dbh + ks * b_k16, idesc, acc);
            acc = 1;
          }
          umma_commit(empty_bar(s));
          if (chunk_left == 0 && i < num_kb) umma_commit(chunk_full_bar);
        }
        __syncwarp();
      }
    }
    if (issuer) umma_commit(tmem_full_bar);
    __syncwarp();
  } else {
    // ===================== workers: warps 2..9 =====================
    const int q = warp & 3;                  // TMEM lane quarter
    const int half = (warp - 2) >> 2;        // column half
    const uint32_t t_main = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(half * CPW);
    float acc[CPW];
#pragma unroll
    for (int j = 0; j < CPW; ++j) acc[j] = 0.f;
    const int n_mid = (num_kb - 1) / H3_CHUNK_KB;           // chunks that end before the last k-block
    const float comp = ph.comp_per_mma * (float)(H3_CHUNK_KB * (BK / UK));
    for (int c = 0; c < n_mid; ++c) {
      mbar_wait(chunk_full_bar, (uint32_t)c & 1u);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < NCH; ++cc) {
        uint32_t v[32];
        tmem_ld32(t_main + (uint32_t)(cc * 32), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[cc * 32 + j] += fmaf(__uint_as_float(v[j]), comp, __uint_as_float(v[j]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(chunk_empty_bar);
    }
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    float sa, sb, ia, ib;
    h3_slot_scale(ph.a_amax, sa, ia);
    h3_slot_scale(ph.b_amax, sb, ib);
    const float inv = ia * ib;
    const float comp_last = ph.comp_per_mma * (float)((num_kb - n_mid * H3_CHUNK_KB) * (BK / UK));
#pragma unroll
    for (int cc = 0; cc < NCH; ++cc) {         // last chunk of the main accumulator + the cross-term accumulator
      uint32_t v[32];
      tmem_ld32(t_main + (uint32_t)(cc * 32), v);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[cc * 32 + j] += fmaf(__uint_as_float(v[j]), comp_last, __uint_as_float(v[j]));
      tmem_ld32(t_main + (uint32_t)(BN + cc * 32), v);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[cc * 32 + j] = (acc[cc * 32 + j] + __uint_as_float(v[j])) * inv;
    }
    // ---- epilogue: this warp's 32 x CPW accumulators -> staging tile -> full-row stores (store_staged)
    float* Cz = p.C + (size_t)blockIdx.z * p.slab_stride;
    const bool vec = epilogue_vec_ok(p, Cz);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    float4* stg = reinterpret_cast<float4*>(base_ptr + (32 * CPW * 4) * (warp - 2));
    const int row = m0 + 32 * q + lane;
    const int cw0 = n0 + half * CPW;
    float vmax = 0.f;
    float* const vm = p.c_amax ? &vmax : nullptr;
    if (cw0 < p.N) {                           // warp-uniform
      if (vec) {
#pragma unroll
        for (int sl = 0; sl < CPW / 4; ++sl) stage_put<CPW>(stg, lane, sl, acc[4 * sl], acc[4 * sl + 1], acc[4 * sl + 2], acc[4 * sl + 3]);
        __syncwarp();
        store_staged<CPW>(p, Cz, stg, lane, m0 + 32 * q, cw0, vm);
      } else if (row < p.M) {
#pragma unroll
        for (int cc = 0; cc < NCH; ++cc) store_row_scalar(p, Cz, row, cw0 + cc * 32, acc + cc * 32, vm);
      }
    }
    if (p.c_amax) {
      const uint32_t mx = __reduce_max_sync(0xffffffffu, __float_as_uint(vmax));
      if (lane == 0 && mx) atomicMax(p.c_amax + 1, mx);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// f16x3, persistent version (the one the big layers run).  ncu on the kernel above (16384 x 1024 x 1024): the tensor
// pipe is busy 57 % of a CTA's life -- the main loop runs at the nominal MMA rate, the rest is the prologue and an
// epilogue nothing overlaps.  Here one CTA per SM walks over tiles (tile = blockIdx.x + j * gridDim.x, n fastest so
// neighbouring SMs share A rows in L2) and the accumulator is double-buffered in TMEM per 256-k CHUNK:
//   * main and cross terms share ONE accumulator of BN columns (3 MMAs per 16 k; the cross terms are 2^-11 of the sum,
//     the drain into fp32 registers every 48 instructions keeps the truncation loss where the kernel above has it);
//   * chunk g lives in buffer g & 1: the MMA warp fills one buffer while the workers drain the other, and the last
//     drain of a tile is followed by its global stores while the tensor core is already on the next tile.
// Warps as above (0 TMA, 1 MMA, 2..9 workers); barriers: full/empty per stage, acc_full/acc_empty per buffer.
// ---------------------------------------------------------------------------------------------------------------
// PAIR: tcgen05 cta_group::2 -- a cluster of two CTAs on one TPC computes a 256 x BN tile; each CTA stages its own 128 rows
// of A and its own HALF of B, the leader's single thread issues M = 256 MMAs that read both shared memories.  Why: the
// 1-CTA main loop pulls 48 KB per 32-k block and SM (62 B/clk/SM, 9.3 KB/clk chip-wide) through an L2 that delivers
// ~6.3 KB/clk to the SMs: measured 33k cycles per 128x256x1024 tile against 24.6k of MMA issue -- L2-bound.  Pairs cut
// the fill to 32 KB per block (and make room for 5 stages instead of 3).
template <int BN, bool SINGLE = false, bool PAIR = false>    // SINGLE: one 16-bit plane per operand and one MMA per 16 k (precision "bf16")
struct CfgP {
  // k-block: 32 elements (64-byte rows, SWIZZLE_64B) with two planes per operand; the one-plane (bf16) kernel takes 64
  // (128-byte rows, SWIZZLE_128B): half as many TMA row requests per byte -- its main loop waited for its loads
  static constexpr int BK = SINGLE ? 64 : 32;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_ROWS = PAIR ? BN / 2 : BN;            // rows of B this CTA stages
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = (SINGLE ? 1 : 2) * (A_BYTES + B_BYTES);
  static constexpr int EPI_COLS = 64;           // columns per staging pass (32: a 4th stage fits, but 128-byte row segments store 40 % slower)
  static constexpr int EPI_BYTES = 8 * 32 * EPI_COLS * 4;           // 8 worker warps x 32 rows x 64 floats
  static constexpr int STAGES = (226 * 1024 - EPI_BYTES) / STAGE_BYTES > 6 ? 6 : (226 * 1024 - EPI_BYTES) / STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr uint32_t K_SBO = 8u * BK * 2u;
  static constexpr uint32_t MN_BOX_BYTES = BK * 128u;
};

// ---- cluster / cta_group::2 helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank0(uint32_t addr) {      // shared::cluster address of `addr` in CTA rank 0
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(addr));
  return r;
}
// Remote arrive without the cluster-scope release (measured in round 1: `arrive.release.cluster` stalls the issuing warp
// for ~1.5k cycles).  What the leader's MMA must observe is ordered by the tcgen05 fences the caller has executed.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  __threadfence_block();
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of a CTA pair: data lands in THIS CTA's shared memory, the transaction bytes are counted on the barrier
// `bar_cluster` (a shared::cluster address: the leader's full barrier)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_f16_2cta(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {   // arrives on `bar` (same offset) in both CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}

struct ParamsP {
  Params p;
  const uint32_t* a_amax; const uint32_t* b_amax;
  float comp_per_mma;
  int tiles_m, tiles_n, total_tiles;
  int chunk_kb;                 // k-blocks per accumulator chunk (drain period)
  int bf16;                     // SINGLE kernel: operands are bf16 (instruction descriptor format 1)
  int repair;                   // second launch of a layer whose output exists only as fp16 planes: exits at once when the
                                //   sticky scale the first launch used fits max|C| (the normal case), else recomputes the
                                //   layer and writes the planes with the scale that max|C| asks for
};

// Warps of the persistent kernel: warpgroup 0 = {TMA producer, MMA issuer, two idle warps}, warpgroups 1 and 2 = the eight
// workers.  Registers are re-split after the set-up (setmaxnreg): an SM sub-partition holds 16K registers and, with ten
// warps, three warps -- a cap of 168 per thread, under which the workers (128 accumulator registers + a 32-register TMEM
// load + addressing) spilled 264 .. 424 bytes in the drain loop.  Twelve warps = three per sub-partition (one of
// warpgroup 0, two workers): 40 + 2 x 232 = 504 <= 512.
constexpr int P_THREADS = 384;
constexpr int P_WORKER0 = 4;          // first worker warp
constexpr int P_REGS_LIGHT = 40, P_REGS_WORKER = 232;
template <int BN, bool SINGLE, bool PAIR>
__global__ void __launch_bounds__(P_THREADS, 1)
gemm_tc_h3p_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                   const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                   const ParamsP pp) {
  using C = CfgP<BN, SINGLE, PAIR>;
  const Params& p = pp.p;
  if (pp.repair) {      // uniform over the grid: two words every thread reads the same
    const uint32_t W = __ldcg(p.c_amax), mx = __ldcg(p.c_amax + 1);
    if (h3_eff_word(W, mx) == W) return;
  }
  constexpr int BK = C::BK, UK = 16;
  constexpr int CPW = BN / 2, NCH = CPW / 32;
  static_assert(CPW % C::EPI_COLS == 0, "BN must be a multiple of 128");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* const base_ptr = smem_raw + (base - raw_u32);
  const uint32_t bars = base + C::STAGES * C::STAGE_BYTES + C::EPI_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
  auto acc_full_bar = [&](int b) { return bars + 8u * (2 * C::STAGES + b); };
  auto acc_empty_bar = [&](int b) { return bars + 8u * (2 * C::STAGES + 2 + b); };
  const uint32_t tmem_ptr_addr = bars + 8u * (2 * C::STAGES + 4);
  auto a_hi = [&](int s) { return base + (uint32_t)s * C::STAGE_BYTES; };
  auto b_hi = [&](int s) { return a_hi(s) + C::A_BYTES; };
  auto a_lo = [&](int s) { return b_hi(s) + C::B_BYTES; };
  auto b_lo = [&](int s) { return a_lo(s) + C::A_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;     // CTA of the pair; rank 0 issues the MMAs
  const bool leader = rank == 0;
  const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;          // tile walker: a CTA or a CTA pair
  const int units = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int TM = PAIR ? 2 * BM : BM;                   // rows of one tile
  const int kb_total = (p.K + BK - 1) / BK;
  const int tiles_mn = pp.tiles_m * pp.tiles_n;
  const long long t_entry = clock64();

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmAh); tma_prefetch_desc(&tmBh);
    if (!SINGLE) { tma_prefetch_desc(&tmAl); tma_prefetch_desc(&tmBl); }
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full_bar(b), 1); mbar_init(acc_empty_bar(b), PAIR ? 16 : 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"((uint32_t)C::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"((uint32_t)C::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr) : "memory");

  // (the register re-split sits INSIDE each warpgroup's branch: ptxas allocates the code behind a setmaxnreg with that
  //  count, and code behind a merge of both branches with the smaller one)
  if (warp < P_WORKER0) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(P_REGS_LIGHT));
  if (warp == 0) {
    // ===================== TMA producer: runs ahead across tile boundaries =====================
    if (lane == 0) {
      uint32_t it = 0;
      // PAIR: every load of both CTAs counts its bytes on the LEADER's full barrier (one expect_tx of 2 stages there)
      auto load = [&](uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
        if (PAIR) tma_load_2d_pair(dst, map, mapa_rank0(bar), c0, c1); else tma_load_2d(dst, map, bar, c0, c1);
      };
      for (int t = unit; t < pp.total_tiles; t += units) {
        const int z = t / tiles_mn, r = t - z * tiles_mn;
        const int m0 = (r / pp.tiles_n) * TM + (int)rank * BM;                 // this CTA's 128 rows of A
        // this CTA's (half of the) B rows; a ragged last n-tile is contracted at its effective width (see the MMA issuer)
        const int nt0 = (r % pp.tiles_n) * BN;
        const int n_eff = min(BN, ((p.N - nt0 + 31) >> 5) << 5);
        const int n0 = nt0 + (PAIR ? (int)rank * (n_eff >> 1) : 0);
        const int kb_begin = z * p.kb_per_split;
        const int num_kb = min(kb_total, kb_begin + p.kb_per_split) - kb_begin;
        for (int i = 0; i < num_kb; ++i, ++it) {
          const int s = (int)(it % C::STAGES);
          const uint32_t phs = (it / C::STAGES) & 1u;
          mbar_wait(empty_bar(s), phs ^ 1u);
          if (!PAIR || leader) mbar_expect_tx(full_bar(s), (PAIR ? 2 : 1) * C::STAGE_BYTES);
          const int k0 = (kb_begin + i) * BK;
          if (!p.a_mn) {
            load(a_hi(s), &tmAh, full_bar(s), k0, m0);
            if (!SINGLE) load(a_lo(s), &tmAl, full_bar(s), k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) {
              load(a_hi(s) + j * C::MN_BOX_BYTES, &tmAh, full_bar(s), m0 + 64 * j, k0);
              if (!SINGLE) load(a_lo(s) + j * C::MN_BOX_BYTES, &tmAl, full_bar(s), m0 + 64 * j, k0);
            }
          }
          if (!p.b_mn) {
            load(b_hi(s), &tmBh, full_bar(s), k0, n0);
            if (!SINGLE) load(b_lo(s), &tmBl, full_bar(s), k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < C::B_ROWS / 64; ++j) {
              load(b_hi(s) + j * C::MN_BOX_BYTES, &tmBh, full_bar(s), n0 + 64 * j, k0);
              if (!SINGLE) load(b_lo(s) + j * C::MN_BOX_BYTES, &tmBl, full_bar(s), n0 + 64 * j, k0);
            }
          }
        }
      }
    }
  } else if (warp == 1 && (!PAIR || leader)) {
    // ===================== MMA issuer (PAIR: the leader CTA's, for both) =====================
    const uint32_t fmt = (SINGLE && pp.bf16) ? ((1u << 7) | (1u << 10)) : 0u;      // A / B format: 0 = fp16, 1 = bf16
    // (N field per tile: a ragged last n-tile -- the 264-column first-layer weight gradient has 8 valid columns in its
    //  second tile -- is contracted with N = the valid width rounded up to 32, not 256: an eighth of the tensor-core work
    //  and energy for that tile.  cta_group::2 takes N/2 columns from each CTA, so the producer loads CTA 1's B rows from
    //  n0 + N/2; accumulator column j is output column n0 + j either way.)
    const uint32_t idesc0 = (1u << 4) | fmt | ((uint32_t)(p.a_mn ? 1 : 0) << 15) | ((uint32_t)(p.b_mn ? 1 : 0) << 16) |
                            ((uint32_t)(TM >> 4) << 24);
    const uint32_t a_lbo = p.a_mn ? C::MN_BOX_BYTES : 16u, b_lbo = p.b_mn ? C::MN_BOX_BYTES : 16u;
    const uint32_t a_sbo = p.a_mn ? 1024u : C::K_SBO, b_sbo = p.b_mn ? 1024u : C::K_SBO;
    constexpr uint32_t K_LAY = BK == 64 ? 2u : 4u;            // K-major tiles: SWIZZLE_128B (128-byte rows) | SWIZZLE_64B
    const uint32_t a_lay = p.a_mn ? 2u : K_LAY, b_lay = p.b_mn ? 2u : K_LAY;
    const uint64_t dA0 = smem_desc(a_hi(0), a_lbo, a_sbo, a_lay);
    const uint64_t dB0 = smem_desc(b_hi(0), b_lbo, b_sbo, b_lay);
    const uint64_t a_k16 = p.a_mn ? (2048u >> 4) : (32u >> 4), b_k16 = p.b_mn ? (2048u >> 4) : (32u >> 4);
    constexpr uint64_t LO16 = (C::A_BYTES + C::B_BYTES) >> 4, STAGE16 = C::STAGE_BYTES >> 4;
    const bool issuer = elect_one();
    uint32_t it = 0, g = 0;
    long long* const dbg = (p.dbg && blockIdx.x == 0) ? p.dbg : nullptr;
    long long w_empty = 0, w_full = 0;
    const long long t_begin = clock64();
    for (int t = unit; t < pp.total_tiles; t += units) {
      const int z = t / tiles_mn;
      const int kb_begin = z * p.kb_per_split;
      const int num_kb = min(kb_total, kb_begin + p.kb_per_split) - kb_begin;
      const int nt0 = ((t - z * tiles_mn) % pp.tiles_n) * BN;
      const uint32_t idesc = idesc0 | ((uint32_t)(min(BN, ((p.N - nt0 + 31) >> 5) << 5) >> 3) << 17);
      for (int kb = 0; kb < num_kb; ++g) {
        const uint32_t b = g & 1u;
        const long long c0 = dbg ? clock64() : 0;
        mbar_wait(acc_empty_bar(b), ((g >> 1) & 1u) ^ 1u);      // the workers have copied chunk g-2 out of this buffer
        if (dbg) w_empty += clock64() - c0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + b * BN;
        const int n = min(pp.chunk_kb, num_kb - kb);
        uint32_t acc = 0;
        for (int j = 0; j < n; ++j, ++it) {
          const int s = (int)(it % C::STAGES);
          const long long c1 = dbg ? clock64() : 0;
          mbar_wait(full_bar(s), (it / C::STAGES) & 1u);
          if (dbg) w_full += clock64() - c1;
          tc_fence_after();
          if (issuer) {
            const uint64_t dah = dA0 + (uint64_t)s * STAGE16, dbh = dB0 + (uint64_t)s * STAGE16;
#pragma unroll
            for (int ks = 0; ks < BK / UK; ++ks) {
              if (PAIR) umma_f16_2cta(d_tmem, dah + ks * a_k16, dbh + ks * b_k16, idesc, acc);
              else umma_f16(d_tmem, dah + ks * a_k16, dbh + ks * b_k16, idesc, acc);
              acc = 1;
              if (!SINGLE) {
                if (PAIR) {
                  umma_f16_2cta(d_tmem, dah + LO16 + ks * a_k16, dbh + ks * b_k16, idesc, 1u);
                  umma_f16_2cta(d_tmem, dah + ks * a_k16, dbh + LO16 + ks * b_k16, idesc, 1u);
                } else {
                  umma_f16(d_tmem, dah + LO16 + ks * a_k16, dbh + ks * b_k16, idesc, 1u);
                  umma_f16(d_tmem, dah + ks * a_k16, dbh + LO16 + ks * b_k16, idesc, 1u);
                }
              }
            }
            if (PAIR) umma_commit_2cta(empty_bar(s)); else umma_commit(empty_bar(s));     // PAIR: frees the stage in both CTAs
          }
          __syncwarp();
        }
        if (issuer) { if (PAIR) umma_commit_2cta(acc_full_bar(b)); else umma_commit(acc_full_bar(b)); }
        __syncwarp();
        kb += n;
      }
    }
    if (dbg && lane == 0) { dbg[0] = clock64() - t_begin; dbg[1] = w_empty; dbg[2] = w_full; dbg[7] = t_begin - t_entry; }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(P_REGS_WORKER));
    // ===================== workers: chunk drains + epilogue =====================
    const int q = warp & 3;
    const int half = (warp - P_WORKER0) >> 2;
    const uint32_t t_lane = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(half * CPW);
    float sa, sb, ia, ib;
    h3_slot_scale(pp.a_amax, sa, ia);
    h3_slot_scale(pp.b_amax, sb, ib);
    const float inv = SINGLE ? 1.0f : ia * ib;       // one-plane operands are not scaled
    float4* stg = reinterpret_cast<float4*>(base_ptr + C::STAGES * C::STAGE_BYTES + (32 * C::EPI_COLS * 4) * (warp - P_WORKER0));
    uint32_t g = 0;
    // worker-decided epilogue fields: the scale of C's planes is read from C's slot; bits_out layers apply bias + ReLU
    // before staging, not in the store loop
    EpiOv ov{p.c_hi, 1.f, p.bias, p.relu};
    const float* const bias_pre = p.bias;
    if (CPW == 128 && p.bits_out) { ov.bias = nullptr; ov.relu = 0; }
    if (ov.c_hi) {
      const uint32_t W = p.c_amax[0];            // prepared by h3_prep_kernel before this launch; 0 = no history
      float is;
      if (pp.repair) h3_scale(h3_eff_word(W, p.c_amax[1]), ov.c_scale, is);     // the scale max|C| asks for (now known)
      else if (W == 0u) ov.c_hi = nullptr;       // the repair pass will write the planes once max|C| is known
      else h3_scale(W, ov.c_scale, is);
    }
    float vmax = 0.f;
    float* const vm = (p.c_amax && !pp.repair) ? &vmax : nullptr;
    long long* const dbg = (p.dbg && blockIdx.x == 0 && warp == P_WORKER0) ? p.dbg : nullptr;
    long long w_accfull = 0, t_drain = 0, t_epi = 0;
    const long long t_begin = clock64();
    uint32_t acc_empty_remote[2];
    if (PAIR) { acc_empty_remote[0] = mapa_rank0(acc_empty_bar(0)); acc_empty_remote[1] = mapa_rank0(acc_empty_bar(1)); }
    for (int t = unit; t < pp.total_tiles; t += units) {
      const int z = t / tiles_mn, r = t - z * tiles_mn;
      const int m0 = (r / pp.tiles_n) * TM + (int)rank * BM, n0 = (r % pp.tiles_n) * BN;     // this CTA's 128 accumulator rows
      const int kb_begin = z * p.kb_per_split;
      const int num_kb = min(kb_total, kb_begin + p.kb_per_split) - kb_begin;
      float acc[CPW];
#pragma unroll
      for (int j = 0; j < CPW; ++j) acc[j] = 0.f;
      // ReLU bit plane of the mask source: this lane's accumulator row x the warp's CPW columns = CPW / 32 words = one
      // 16-byte load at the start of the epilogue (host: CPW == 128, ld_bits % 4 == 0, N % 128 == 0).  Only a prefetch
      // here: holding the four words across the chunk drains cost more (register cap 168: 29k -> 47k cycles of drains
      // per four tiles) than the L2 round trip they would hide.
      if (CPW == 128 && p.bits_in) {
        const int brow = m0 + 32 * q + lane, bc0 = n0 + half * CPW;
        if (brow < p.M && bc0 < p.N)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p.bits_in + (size_t)brow * p.ld_bits + (bc0 >> 5)));
      }
      for (int kb = 0; kb < num_kb; ++g) {
        const uint32_t b = g & 1u;
        const int n = min(pp.chunk_kb, num_kb - kb);
        const float comp = pp.comp_per_mma * (float)(n * (BK / UK) * (SINGLE ? 1 : 3));
        const long long c0 = dbg ? clock64() : 0;
        mbar_wait(acc_full_bar(b), (g >> 1) & 1u);
        const long long c1 = dbg ? clock64() : 0;
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < NCH; ++cc) {
          uint32_t v[32];
          tmem_ld32(t_lane + b * BN + (uint32_t)(cc * 32), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[cc * 32 + j] += fmaf(__uint_as_float(v[j]), comp, __uint_as_float(v[j]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {       // PAIR: the leader's MMA waits for the workers of BOTH CTAs (16 arrivals on its barrier)
          if (PAIR && !leader) mbar_arrive_cluster_relaxed(b ? acc_empty_remote[1] : acc_empty_remote[0]);
          else mbar_arrive(acc_empty_bar(b));
        }
        if (dbg) { w_accfull += c1 - c0; t_drain += clock64() - c1; }
        kb += n;
      }
      const long long c2 = dbg ? clock64() : 0;
      // ---- epilogue: 32 rows x CPW columns of this warp
      float* Cz = p.C ? p.C + (size_t)z * p.slab_stride : nullptr;
      const bool vec = epilogue_vec_ok(p, Cz);
      const int row = m0 + 32 * q + lane;
      const int cw0 = n0 + half * CPW;
      uint4 mbits = make_uint4(0u, 0u, 0u, 0u), obits = make_uint4(0u, 0u, 0u, 0u);
      if (CPW == 128 && p.bits_in && row < p.M && cw0 < p.N)
        mbits = __ldg(reinterpret_cast<const uint4*>(p.bits_in + (size_t)row * p.ld_bits + (cw0 >> 5)));
      {
#pragma unroll
      for (int ps = 0; ps < CPW / C::EPI_COLS; ++ps) {
        const int c0 = cw0 + ps * C::EPI_COLS;
        // warp-uniform.  A warp whose 32 rows all lie beyond M has nothing to store:
        // with M = 16385 (the discriminator chain) the guarded edge path below ran for all 16 warps of the last row tile
        // and made every such layer ~20 us longer than its 16384-row twin.
        if (c0 < p.N && m0 + 32 * q < p.M) {
          if (vec) {
            if (CPW == 128 && p.bits_in) {      // the ReLU mask, applied where a lane still owns its row
              static_assert(C::EPI_COLS == 64, "the bit-plane layout is defined per 64-column staging pass");
              const uint32_t w0 = ps == 0 ? mbits.x : mbits.z, w1 = ps == 0 ? mbits.y : mbits.w;      // (layout: see "ReLU bit planes" above)
#pragma unroll
              for (int sl = 0; sl < C::EPI_COLS / 4; ++sl) {
                const int j = ps * C::EPI_COLS + 4 * sl;
                stage_put<C::EPI_COLS>(stg, lane, sl, ((w0 >> sl) & 1u) ? acc[j] * inv : 0.f, ((w0 >> (16 + sl)) & 1u) ? acc[j + 1] * inv : 0.f,
                                       ((w1 >> sl) & 1u) ? acc[j + 2] * inv : 0.f, ((w1 >> (16 + sl)) & 1u) ? acc[j + 3] * inv : 0.f);
              }
            } else if (CPW == 128 && p.bits_out) {   // ReLU layer that leaves its mask behind: bias + clamp + bits here
              uint32_t w0 = 0u, w1 = 0u;
#pragma unroll
              for (int sl = 0; sl < C::EPI_COLS / 4; ++sl) {
                const int j = ps * C::EPI_COLS + 4 * sl;
                float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                if (bias_pre) bb = __ldg(reinterpret_cast<const float4*>(bias_pre + c0 + 4 * sl));      // one address per warp
                // (same operations, in the same order, as the store loop: acc * inv is exact -- inv is a power of two)
                const float a0 = fmaxf(__fadd_rn(__fmul_rn(acc[j], inv), bb.x), 0.f), a1 = fmaxf(__fadd_rn(__fmul_rn(acc[j + 1], inv), bb.y), 0.f);
                const float a2 = fmaxf(__fadd_rn(__fmul_rn(acc[j + 2], inv), bb.z), 0.f), a3 = fmaxf(__fadd_rn(__fmul_rn(acc[j + 3], inv), bb.w), 0.f);
                if (a0 > 0.f) w0 |= 1u << sl;
                if (a1 > 0.f) w0 |= 1u << (16 + sl);
                if (a2 > 0.f) w1 |= 1u << sl;
                if (a3 > 0.f) w1 |= 1u << (16 + sl);
                stage_put<C::EPI_COLS>(stg, lane, sl, a0, a1, a2, a3);
              }
              if (ps == 0) { obits.x = w0; obits.y = w1; } else { obits.z = w0; obits.w = w1; }
            } else {
#pragma unroll
            for (int sl = 0; sl < C::EPI_COLS / 4; ++sl) {
              const int j = ps * C::EPI_COLS + 4 * sl;
              stage_put<C::EPI_COLS>(stg, lane, sl, acc[j] * inv, acc[j + 1] * inv, acc[j + 2] * inv, acc[j + 3] * inv);
            }
            }
            __syncwarp();
            if (m0 + 32 * q + 32 <= p.M && c0 + C::EPI_COLS <= p.N && !p.accumulate) {     // warp-uniform
              if (SINGLE && p.C16) {
                if (p.mask || p.mask16) store_staged_interior<C::EPI_COLS, true, false, true>(p, ov, Cz, stg, lane, m0 + 32 * q, c0, vm);
                else store_staged_interior<C::EPI_COLS, false, false, true>(p, ov, Cz, stg, lane, m0 + 32 * q, c0, vm);
              // (a variant with eight columns per lane -- 16-byte plane stores, 16-byte mask loads -- measured slower in the
              //  whole update: 90.8 against 88.4 ms per iteration on one box)
              } else if (!SINGLE && ov.c_hi) {
                if (p.mask || p.mask16) store_staged_interior<C::EPI_COLS, true, true>(p, ov, Cz, stg, lane, m0 + 32 * q, c0, vm);
                else store_staged_interior<C::EPI_COLS, false, true>(p, ov, Cz, stg, lane, m0 + 32 * q, c0, vm);
              } else {
                if (p.mask || p.mask16) store_staged_interior<C::EPI_COLS, true, false>(p, ov, Cz, stg, lane, m0 + 32 * q, c0, vm);
                else store_staged_interior<C::EPI_COLS, false, false>(p, ov, Cz, stg, lane, m0 + 32 * q, c0, vm);
              }
            } else if (c0 + C::EPI_COLS <= p.N && !p.accumulate) {      // ragged rows only
              store_staged_rows<C::EPI_COLS>(p, ov, Cz, stg, lane, m0 + 32 * q, c0, p.M - (m0 + 32 * q), vm);
            } else {
              store_staged<C::EPI_COLS>(p, ov, Cz, stg, lane, m0 + 32 * q, c0, vm);
            }
            __syncwarp();
          } else if (row < p.M) {
            float f[32];
#pragma unroll
            for (int cc = 0; cc < C::EPI_COLS / 32; ++cc) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = acc[ps * C::EPI_COLS + cc * 32 + j] * inv;
              store_row_scalar(p, ov, Cz, row, c0 + cc * 32, f, vm);
            }
          }
        }
      }
      }
      if (CPW == 128 && p.bits_out && row < p.M && cw0 < p.N)
        *reinterpret_cast<uint4*>(p.bits_out + (size_t)row * p.ld_bits + (cw0 >> 5)) = obits;
      if (dbg) t_epi += clock64() - c2;
    }
    if (dbg && lane == 0) { dbg[3] = clock64() - t_begin; dbg[4] = w_accfull; dbg[5] = t_drain; dbg[6] = t_epi; dbg[8] = clock64() - t_entry; }
    if (p.c_amax && !pp.repair) {
      const uint32_t mx = __reduce_max_sync(0xffffffffu, __float_as_uint(vmax));
      if (lane == 0 && mx) atomicMax(p.c_amax + 1, mx);
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
  }
}

// 2-D fp16 tensor map: memory [outer, inner] with `ld` elements between rows; K-major: box {32 k, box_rows}, 64-byte
// swizzle; MN-major: box {64 mn, 32 k}, 128-byte swizzle.
static bool make_map_f16(CUtensorMap* map, const void* ptr, long long inner, long long outer, long long ld, int box_inner,
                         int box_rows, bool bf16 = false) {
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = g_encode(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims,
                        strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, box_inner == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BN>
static int launch_h3(cudaStream_t st, const CUtensorMap& tah, const CUtensorMap& tal, const CUtensorMap& tbh,
                     const CUtensorMap& tbl, const ParamsH3& p, dim3 grid) {
  using C = CfgH3<BN>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_tc_h3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess) {
      addk_set_error("gemm_tc: cannot raise the dynamic shared memory limit");
      return ADDK_ERR_LAUNCH;
    }
    configured = true;
  }
  gemm_tc_h3_kernel<BN><<<grid, X3_THREADS, C::SMEM_BYTES, st>>>(tah, tal, tbh, tbl, p);
  return ADDK_OK;
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int BN, bool SINGLE = false, bool PAIR = false>
static int launch_h3p(cudaStream_t st, const CUtensorMap& tah, const CUtensorMap& tal, const CUtensorMap& tbh,
                      const CUtensorMap& tbl, ParamsP& pp, int M, int N, int split) {
  using C = CfgP<BN, SINGLE, PAIR>;
  pp.repair = 0;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_tc_h3p_kernel<BN, SINGLE, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess) {
      addk_set_error("gemm_tc: cannot raise the dynamic shared memory limit");
      return ADDK_ERR_LAUNCH;
    }
    configured = true;
  }
  constexpr int TM = PAIR ? 2 * BM : BM;
  pp.tiles_m = (M + TM - 1) / TM; pp.tiles_n = (N + BN - 1) / BN; pp.total_tiles = pp.tiles_m * pp.tiles_n * split;
  // (74 / 99 / 128 CTAs per layer, so that layers of the three streams run side by side, measured the same 2.6 ms per
  // optimizer step as one CTA per SM)
  const int launches = (pp.p.no_f32 && pp.p.c_hi) ? 2 : 1;      // planes-only output: main launch + repair launch
  if (!PAIR) {
    const int grid = pp.total_tiles < sm_count() ? pp.total_tiles : sm_count();
    for (int l = 0; l < launches; ++l) {
      pp.repair = l;
      gemm_tc_h3p_kernel<BN, SINGLE, PAIR><<<l ? 1 : grid, P_THREADS, C::SMEM_BYTES, st>>>(tah, tal, tbh, tbl, pp);
    }
    return ADDK_OK;
  }
  const int pairs = sm_count() / 2;
  const int units = pp.total_tiles < pairs ? pp.total_tiles : pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * units, 1, 1);
  cfg.blockDim = dim3(P_THREADS, 1, 1);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  for (int l = 0; l < launches; ++l) {
    pp.repair = l;
    // the repair launch (exits at once unless the sticky scale missed max|C|): ONE CTA pair walks every tile -- the rare
    // repair takes milliseconds instead of ~90 us, the common early exit does not have to place 148 CTAs
    if (l) cfg.gridDim = dim3(2, 1, 1);
    if (cudaLaunchKernelEx(&cfg, gemm_tc_h3p_kernel<BN, SINGLE, PAIR>, tah, tal, tbh, tbl, (const ParamsP)pp) != cudaSuccess) {
      addk_set_error("gemm_tc: cluster launch of the CTA-pair kernel failed");
      return ADDK_ERR_LAUNCH;
    }
  }
  return ADDK_OK;
}

// ---- the split pre-pass: max|x| of a [rows, cols] fp32 tensor (pitch ld), then the two fp16 planes ----------------
__global__ void h3_amax_kernel(const float* __restrict__ x, long long rows, int cols, int ld, uint32_t* __restrict__ slot) {
  uint32_t m = 0;
  const bool flat = ld == cols;
  const bool vec = ((cols & 3) == 0) && ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (vec) {
    const int c4 = cols >> 2;
    if (flat) {
      const long long n4 = rows * c4;
      const long long step = (long long)gridDim.x * blockDim.x;
      for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 4 * step) {     // 4 loads in flight
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = i0 + u * step < n4 ? reinterpret_cast<const float4*>(x)[i0 + u * step] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          m = max(max(m, __float_as_uint(fabsf(v[u].x))), max(__float_as_uint(fabsf(v[u].y)), max(__float_as_uint(fabsf(v[u].z)), __float_as_uint(fabsf(v[u].w)))));
      }
    } else {
      for (long long r = blockIdx.x; r < rows; r += gridDim.x)
        for (int c = threadIdx.x; c < c4; c += blockDim.x) {
          const float4 v = *reinterpret_cast<const float4*>(x + r * ld + 4 * c);
          m = max(max(m, __float_as_uint(fabsf(v.x))), max(__float_as_uint(fabsf(v.y)), max(__float_as_uint(fabsf(v.z)), __float_as_uint(fabsf(v.w)))));
        }
    }
  } else {
    for (long long r = blockIdx.x; r < rows; r += gridDim.x)
      for (int c = threadIdx.x; c < cols; c += blockDim.x) m = max(m, __float_as_uint(fabsf(x[r * ld + c])));
  }
  m = __reduce_max_sync(0xffffffffu, m);
  __shared__ uint32_t sm[32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0u;
    m = __reduce_max_sync(0xffffffffu, m);
    if (threadIdx.x == 0 && m) atomicMax(slot + 1, m);
  }
}

__device__ __forceinline__ void h3_split1(float x, float s, uint16_t& hi, uint16_t& lo) {
  const float xs = x * s;                                   // exact (power of two)
  const __half h = __float2half_rn(xs);
  const float r = xs - __half2float(h);                     // exact difference
  hi = __half_as_ushort(h);
  lo = __half_as_ushort(__float2half_rn(r));
}

// slot[0] = effective sticky word, slot[1] = 0: issued before a dense layer that writes the planes of its output
__global__ void h3_prep_kernel(uint32_t* slot, int keep) {
  slot[0] = keep ? h3_eff_word(slot[0], slot[1]) : 0u;      // keep = 0: no history, the scale follows max|x| alone
  slot[1] = 0u;
}

// The same for a whole table of slots in one launch (an entry point's slots, before its first layer): slots
// [0, n_sticky) keep their history as above, slots [n_sticky, n_total) are cleared (the once-per-call slots).
__global__ void h3_prep_all_kernel(uint32_t* slots, int n_sticky, int n_total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_total) return;
  slots[2 * i] = i < n_sticky ? h3_eff_word(slots[2 * i], slots[2 * i + 1]) : 0u;
  slots[2 * i + 1] = 0u;
}

// colpart (optional, flat vectorised tensors whose float4 columns divide the block: cols / 4 in {256, 128, 64, 32}):
// the pass also leaves per-block column sums of x in colpart[blockIdx.x, 0 .. cols) -- every thread only ever sees one
// float4 column (the grid stride is a multiple of cols / 4), so the bias gradient 1^T dY costs no extra read of dY.
__global__ void __launch_bounds__(256)
h3_split_kernel(const float* __restrict__ x, long long rows, int cols, int ld, const uint32_t* __restrict__ slot,
                uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int repair, float* __restrict__ colpart) {
  __shared__ float4 s_cs[256];
  float s, inv;
  h3_slot_scale(slot, s, inv);
  if (repair && h3_eff_word(slot[0], slot[1]) == slot[0]) return;      // the epilogue's planes stand
  const bool flat = ld == cols;
  const bool vec = ((cols & 3) == 0) && ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(hi) & 7) == 0) && ((reinterpret_cast<uintptr_t>(lo) & 7) == 0);
  if (vec) {
    const int c4 = cols >> 2;
    const long long nrow = flat ? 1 : rows;
    const long long per = flat ? rows * c4 : c4;
    for (long long r = flat ? 0 : blockIdx.x; r < nrow; r += gridDim.x) {
      const long long start = flat ? (long long)blockIdx.x * blockDim.x + threadIdx.x : threadIdx.x;
      const long long step = flat ? (long long)gridDim.x * blockDim.x : blockDim.x;
      // back to front: the end of the tensor is what the producer (or the max pass) touched last and is still in L2;
      // four 128-bit loads in flight per thread
      float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
      for (long long c0 = start; c0 < per; c0 += 4 * step) {
        float4 v[4];
        long long off[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long long c = c0 + u * step;
          off[u] = c < per ? r * ld + 4 * (per - 1 - c) : -1;
          if (off[u] >= 0) v[u] = *reinterpret_cast<const float4*>(x + off[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (off[u] < 0) continue;
          uint2 h2, l2;
          h3_split4(v[u], s, h2, l2);
          *reinterpret_cast<uint2*>(hi + off[u]) = h2;
          *reinterpret_cast<uint2*>(lo + off[u]) = l2;
          cs.x += v[u].x; cs.y += v[u].y; cs.z += v[u].z; cs.w += v[u].w;
        }
      }
      if (colpart && flat) {                     // host guarantees: 256 % c4 == 0, grid stride % c4 == 0
        s_cs[threadIdx.x] = cs;
        __syncthreads();
        if ((int)threadIdx.x < c4) {
          float4 t = s_cs[threadIdx.x];
          for (int j = threadIdx.x + c4; j < 256; j += c4) { t.x += s_cs[j].x; t.y += s_cs[j].y; t.z += s_cs[j].z; t.w += s_cs[j].w; }
          // this thread's float4 column: element (per - 1 - c) with c = tid (mod c4)  ->  column group c4 - 1 - tid
          *reinterpret_cast<float4*>(colpart + (size_t)blockIdx.x * cols + 4 * (c4 - 1 - (int)threadIdx.x)) = t;
        }
      }
    }
  } else {
    for (long long r = blockIdx.x; r < rows; r += gridDim.x)
      for (int c = threadIdx.x; c < cols; c += blockDim.x) h3_split1(x[r * ld + c], s, hi[r * ld + c], lo[r * ld + c]);
  }
}

static int h3_convert(cudaStream_t st, const float* x, long long rows, int cols, int ld, void* hi, long long plane, uint32_t* slot,
                      bool have_amax = false, float* colpart = nullptr, int* colpart_rows = nullptr) {
  if (!x || !hi || !slot || rows <= 0 || cols <= 0 || ld < cols || plane <= 0) { addk_set_error("f16x3 convert: bad arguments"); return ADDK_ERR_ARG; }
  // A few alignment-padding columns (29 -> 32 actions, 264 -> 272 observations) are part of the tensor's own
  // buffer and hold zeros: scan them too, flat and vectorised, instead of one block per 29-column row (18 -> 3 us)
  if (cols < ld && ld - cols <= 8 && (ld & 3) == 0) cols = ld;
  const long long work = ld == cols ? (rows * cols / 4 + 255) / 256 : rows;
  const unsigned blocks = (unsigned)(work < 1 ? 1 : (work > 148 * 8 ? 148 * 8 : work));
  if (colpart) {         // column sums ride along only for flat tensors whose float4 columns divide the 256-thread block
    const int c4 = cols / 4;
    const bool okc = ld == cols && (cols & 3) == 0 && c4 >= 32 && c4 <= 256 && 256 % c4 == 0 &&
                     ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(colpart) & 15) == 0);
    if (!okc) colpart = nullptr;
    if (colpart_rows) *colpart_rows = colpart ? (int)blocks : 0;
  }
  if (!have_amax) {      // otherwise the producing dense layer left max|x| in the slot (addk_gemm_args::c_amax)
    if (cudaMemsetAsync(slot, 0, 2 * sizeof(uint32_t), st) != cudaSuccess) { addk_set_error("f16x3 convert: memset failed"); return ADDK_ERR_LAUNCH; }
    h3_amax_kernel<<<blocks, 256, 0, st>>>(x, rows, cols, ld, slot);
  }
  h3_split_kernel<<<blocks, 256, 0, st>>>(x, rows, cols, ld, slot, reinterpret_cast<uint16_t*>(hi), reinterpret_cast<uint16_t*>(hi) + plane, 0, colpart);
  return ADDK_OK;
}

#ifdef ADDK_LEGACY_KERNELS
#include "gemm_tc_legacy.cuh"
#endif

}  // namespace addk_tc

// ids reported by addk_debug_last_gemm_kernel(): which kernel a call was dispatched to
enum { ADDK_K_SGEMM = 0, ADDK_K_TF32 = 10, ADDK_K_TF32X3 = 11, ADDK_K_TF32X3_PAIR = 12, ADDK_K_BF16_TILE = 20,
       ADDK_K_BF16_PERSISTENT = 21, ADDK_K_BF16_PAIR = 22, ADDK_K_H3_TILE = 30, ADDK_K_H3_PERSISTENT = 31, ADDK_K_H3_PAIR = 32 };

// precision: 1 = tf32x3, 2 = tf32.  Shapes the tensor-core tiles do not cover (heads with 1 or 29 outputs,
// contraction shorter than one k-block, misaligned leading dimensions, fused input normalisation) run on the
// exact-fp32 CUDA-core kernel, which is at least as accurate.
__global__ void f32_to_bf16_rows_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int M, int N, int ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * N) return;
  const int r = (int)(i / N), c = (int)(i - (long long)r * N);
  dst[(size_t)r * ld + c] = (uint16_t)(addk_tc::pack_bf16x2(src[(size_t)r * ld + c], 0.f) & 0xFFFFu);
}

// ReLU masks as bit planes (persistent kernels, f16x3 and bf16): validate and hand them to the kernel parameters
static int apply_relu_bits(const addk_gemm_args& a, addk_tc::Params& p, int split) {
  if (!a.relu_bits_out && !a.relu_bits_in) return ADDK_OK;
  const uintptr_t bo = reinterpret_cast<uintptr_t>(a.relu_bits_out), bi = reinterpret_cast<uintptr_t>(a.relu_bits_in);
  if ((a.N & 127) || (a.ld_bits & 3) || a.ld_bits * 32 < a.N || ((bo | bi) & 15) || (a.ldc & 3) || split != 1 || a.accumulate ||
      (reinterpret_cast<uintptr_t>(a.C) & 15) || (reinterpret_cast<uintptr_t>(a.bias) & 15) ||
      (a.relu_bits_out && (!a.relu || a.relu_bits_in || a.relu_mask_src || a.relu_mask_src16))) {
    addk_set_error("gemm: relu_bits_* need N % 128 == 0, ld_bits % 4 == 0 (>= N / 32), 16-byte aligned bit planes, one slab; "
                   "relu_bits_out a ReLU layer without a mask of its own");
    return ADDK_ERR_ARG;
  }
  p.bits_out = a.relu_bits_out; p.ld_bits = a.ld_bits;
  if (a.relu_bits_in) { p.bits_in = a.relu_bits_in; p.mask = nullptr; p.mask16 = nullptr; }
  return ADDK_OK;
}
// column sums per 32-row block (persistent kernels): only the vectorised store paths produce them
static int apply_colsum_partials(const addk_gemm_args& a, addk_tc::Params& p, int split) {
  if (!a.colsum_partials) return ADDK_OK;
  if (split != 1 || a.accumulate || (a.N & 63) || (a.ldc & 3) || (reinterpret_cast<uintptr_t>(a.colsum_partials) & 15) ||
      (reinterpret_cast<uintptr_t>(a.C) & 15) || (reinterpret_cast<uintptr_t>(a.bias) & 15) ||
      (a.relu_mask_src && ((a.ld_mask & 3) || (reinterpret_cast<uintptr_t>(a.relu_mask_src) & 15))) ||
      (a.relu_mask_src16 && ((a.ld_mask & 3) || (reinterpret_cast<uintptr_t>(a.relu_mask_src16) & 7)))) {
    addk_set_error("gemm: colsum_partials need one slab, no accumulate, N % 64 == 0 and 16-byte aligned rows");
    return ADDK_ERR_ARG;
  }
  p.colpart = a.colsum_partials;
  return ADDK_OK;
}

static int gemm_bf16(cudaStream_t st, const addk_gemm_args& a) {
  using namespace addk_tc;
  const int BKh = 64;
  int split = a.split_k > 1 ? a.split_k : 1;
  const int kb_total = (a.K + BKh - 1) / BKh;
  const int kb_per = (kb_total + split - 1) / split;
  if ((long long)kb_per * (split - 1) >= kb_total) return -1;
  Params p;
  p.C = a.C; p.ldc = a.ldc; p.M = a.M; p.N = a.N; p.K = a.K; p.bias = a.bias; p.mask = a.relu_mask_src;
  p.ld_mask = a.ld_mask; p.relu = a.relu; p.accumulate = 0; p.kb_per_split = kb_per;
  p.slab_stride = a.slab_stride > 0 ? a.slab_stride : (long long)a.M * a.ldc;
  p.pair_flags = 0; p.dbg = g_addk_stamps; p.C16 = a.C16; p.c_amax = nullptr; p.c_hi = nullptr; p.c_plane = 0; p.c_scale = 1.f; p.c16_in_staged = 0; p.no_f32 = 0; p.mask16 = nullptr;
  p.a_mn = a.trans_a ? 1 : 0;
  p.b_mn = a.trans_b ? 0 : 1;
  const int BN = a.N > 128 ? 256 : (a.N > 64 ? 128 : 64);
  const int persistent = addk_switches().bf16_persistent;
  if ((a.no_f32 || a.relu_mask_src16) && !(BN == 256 && persistent && split == 1 && a.C16)) {
    addk_set_error("gemm bf16: no_f32 / relu_mask_src16 need the persistent kernel (N > 128, one slab) and a bf16 output");
    return ADDK_ERR_ARG;
  }
  // (split-K weight gradients too since the 64-k blocks and the CTA pairs: 57 -> see profiles/r02_SUMMARY.md)
  if (BN == 256 && persistent && (a.C || a.no_f32) && (split == 1 || a.M > BM)) {     // (split-K weight gradients: the one-tile-per-CTA kernel is faster, 54 vs 62 us)
    // the persistent kernel of the f16x3 mode with one plane per operand: 32-k blocks, 6 stages, the accumulator of a
    // whole tile is one chunk (no precision drains), the epilogue of a tile overlaps the next tile's MMAs
    const int kbt = (a.K + 63) / 64;                 // the one-plane persistent kernel walks 64-k blocks
    const int kbp = (kbt + split - 1) / split;
    if ((long long)kbp * (split - 1) < kbt) {
      ParamsP pp;
      pp.p = p; pp.p.kb_per_split = kbp; pp.p.c16_in_staged = 1;
      pp.p.no_f32 = a.no_f32 ? 1 : 0;
      if (a.relu_mask_src16) { pp.p.mask16 = reinterpret_cast<const uint16_t*>(a.relu_mask_src16); pp.p.mask = nullptr; }
      { const int rcb = apply_relu_bits(a, pp.p, split); if (rcb != ADDK_OK) return rcb; }
      { const int rcc = apply_colsum_partials(a, pp.p, split); if (rcc != ADDK_OK) return rcc; }
      pp.a_amax = nullptr; pp.b_amax = nullptr; pp.comp_per_mma = 0.f; pp.chunk_kb = kbp; pp.bf16 = 1;
      CUtensorMap tah, tbh;
      bool okp = p.a_mn ? make_map_f16(&tah, a.A16, a.M, a.K, a.lda, 64, 64, true) : make_map_f16(&tah, a.A16, a.K, a.M, a.lda, 64, BM, true);
      okp = okp && (p.b_mn ? make_map_f16(&tbh, a.B16, a.N, a.K, a.ldb, 64, 64, true) : make_map_f16(&tbh, a.B16, a.K, a.N, a.ldb, 64, 256, true));
      if (okp && addk_switches().h3_pair && a.M > BM) {
        CUtensorMap tbp = tbh;
        const bool okb = p.b_mn || make_map_f16(&tbp, a.B16, a.K, a.N, a.ldb, 64, 128, true);
        if (okb) { g_addk_last_gemm_kernel = ADDK_K_BF16_PAIR; return launch_h3p<256, true, true>(st, tah, tah, tbp, tbp, pp, a.M, a.N, split); }
      }
      if (okp) { g_addk_last_gemm_kernel = ADDK_K_BF16_PERSISTENT; return launch_h3p<256, true>(st, tah, tah, tbh, tbh, pp, a.M, a.N, split); }
    }
  }
  if (a.relu_bits_in || a.relu_bits_out || a.colsum_partials) { addk_set_error("gemm bf16: relu_bits_* / colsum_partials need the persistent kernel"); return ADDK_ERR_ARG; }
  CUtensorMap ta, tb;
  bool ok = p.a_mn ? make_map_bf16(&ta, a.A16, a.M, a.K, a.lda, 64) : make_map_bf16(&ta, a.A16, a.K, a.M, a.lda, BM);
  ok = ok && (p.b_mn ? make_map_bf16(&tb, a.B16, a.N, a.K, a.ldb, 64) : make_map_bf16(&tb, a.B16, a.K, a.N, a.ldb, BN));
  if (!ok) return -1;
  dim3 grid((a.N + BN - 1) / BN, (a.M + BM - 1) / BM, split);
  g_addk_last_gemm_kernel = ADDK_K_BF16_TILE;
  if (BN == 256) return launch_bf16<256>(st, ta, tb, p, grid);
  if (BN == 128) return launch_bf16<128>(st, ta, tb, p, grid);
  return launch_bf16<64>(st, ta, tb, p, grid);
}

extern "C" int addk_f16x3_convert(void* stream, const float* x, long long rows, int cols, int ld, void* hi16, long long plane,
                                  uint32_t* amax_slot) {
  const int rc = addk_tc::h3_convert((cudaStream_t)stream, x, rows, cols, ld, hi16, plane, amax_slot);
  if (rc != ADDK_OK) return rc;
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}

extern "C" int addk_f16x3_split(void* stream, const float* x, long long rows, int cols, int ld, void* hi16, long long plane,
                                uint32_t* amax_slot, float* colsum_partials, int* colsum_partial_rows) {
  if (colsum_partial_rows) *colsum_partial_rows = 0;
  const int rc = addk_tc::h3_convert((cudaStream_t)stream, x, rows, cols, ld, hi16, plane, amax_slot, true, colsum_partials,
                                     colsum_partial_rows);
  if (rc != ADDK_OK) return rc;
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}
extern "C" int addk_f16x3_prep(void* stream, uint32_t* slot, int keep_sticky_word) {
  if (!slot) return ADDK_ERR_ARG;
  addk_tc::h3_prep_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(slot, keep_sticky_word);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}
extern "C" int addk_f16x3_prep_all(void* stream, uint32_t* slots, int n_sticky, int n_total) {
  if (!slots || n_total <= 0 || n_sticky < 0 || n_sticky > n_total) return ADDK_ERR_ARG;
  addk_tc::h3_prep_all_kernel<<<(n_total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(slots, n_sticky, n_total);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}
extern "C" int addk_f16x3_repair(void* stream, const float* x, long long rows, int cols, int ld, void* hi16, long long plane,
                                 uint32_t* slot) {
  if (!x || !hi16 || !slot || rows <= 0 || cols <= 0 || ld < cols || plane <= 0) return ADDK_ERR_ARG;
  const long long work = ld == cols ? (rows * cols / 4 + 255) / 256 : rows;
  const unsigned blocks = (unsigned)(work < 1 ? 1 : (work > 148 * 8 ? 148 * 8 : work));
  addk_tc::h3_split_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, cols, ld, slot, reinterpret_cast<uint16_t*>(hi16),
                                                                     reinterpret_cast<uint16_t*>(hi16) + plane, 1, nullptr);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}

static float addk_h3_comp() { return addk_switches().h3_comp; }

// precision "f16x3": can this shape run on the fp16 planes (given 16-byte aligned twins)?  Shared with csrc/mlp.cu, whose
// twin bookkeeping must know whether a call is going to convert its operands.
bool addk_gemm_h3_usable(const addk_gemm_args& a) {
  const int split = a.split_k > 1 ? a.split_k : 1;
  const int kb_total = (a.K + 31) / 32;
  const int kb_per = (kb_total + split - 1) / split;
  return ((a.lda & 7) == 0) && ((a.ldb & 7) == 0) && !a.a_mean && a.M >= 16 && a.N >= 16 && a.K >= 16 &&
         (split == 1 || !(a.bias || a.relu || a.relu_mask_src || a.accumulate)) &&
         (long long)kb_per * (split - 1) < kb_total && addk_tc::resolve_encode();
}

static int gemm_h3(cudaStream_t st, const addk_gemm_args& a) {
  using namespace addk_tc;
  const int BKh = 32;
  int split = a.split_k > 1 ? a.split_k : 1;
  const int kb_total = (a.K + BKh - 1) / BKh;
  const int kb_per = (kb_total + split - 1) / split;
  // operand geometry as stored: A is [M,K] (trans_a = 0) or [K,M]; B is [N,K] (trans_b = 1) or [K,N]
  const long long a_rows = a.trans_a ? a.K : a.M, b_rows = a.trans_b ? a.N : a.K;
  const int a_cols = a.trans_a ? a.M : a.K, b_cols = a.trans_b ? a.K : a.N;
  if (a.a16_ready != 1) { const int rc = h3_convert(st, a.A, a_rows, a_cols, a.lda, const_cast<void*>(a.A16), a.a16_plane, a.a_amax, a.a16_ready == 2); if (rc != ADDK_OK) return rc; }
  if (a.b16_ready != 1) { const int rc = h3_convert(st, a.B, b_rows, b_cols, a.ldb, const_cast<void*>(a.B16), a.b16_plane, a.b_amax, a.b16_ready == 2); if (rc != ADDK_OK) return rc; }
  ParamsH3 ph;
  Params& p = ph.p;
  p.C = a.C; p.ldc = a.ldc; p.M = a.M; p.N = a.N; p.K = a.K; p.bias = a.bias; p.mask = a.relu_mask_src;
  p.ld_mask = a.ld_mask; p.relu = a.relu; p.accumulate = a.accumulate; p.kb_per_split = kb_per;
  p.slab_stride = a.slab_stride > 0 ? a.slab_stride : (long long)a.M * a.ldc;
  p.pair_flags = 0;
  p.dbg = g_addk_stamps; p.C16 = nullptr; p.c_amax = split == 1 ? a.c_amax : nullptr;
  p.c_hi = nullptr; p.c_plane = 0; p.c_scale = 1.f; p.c16_in_staged = 0; p.no_f32 = 0; p.mask16 = nullptr;
  p.a_mn = a.trans_a ? 1 : 0;
  p.b_mn = a.trans_b ? 0 : 1;
  ph.a_amax = a.a_amax; ph.b_amax = a.b_amax; ph.comp_per_mma = addk_h3_comp();
  int BN = a.N > 128 ? 256 : (a.N > 64 ? 128 : 64);
  // few, long tiles (the weight gradient of a head: 29 x 512 outputs, K = 16384 / 9): narrower tiles put 4x the SMs to work
  const bool want_planes = a.C16 && a.c16_plane > 0;       // only the persistent 256-wide kernel writes C's planes
  if (want_planes && (BN != 256 || split != 1 || !a.c_amax)) { addk_set_error("gemm f16x3: c16_plane needs N > 128, one slab and c_amax"); return ADDK_ERR_ARG; }
  // (128-wide layers with at most half an SM-round of tiles -- the 1024 x 114 first-layer weight gradient of the
  //  discriminator: 72 CTAs, load-latency bound -- also take 64-wide tiles: twice the CTAs)
  if (!want_planes && BN > 64 && (long long)((a.M + BM - 1) / BM) * ((a.N + BN - 1) / BN) * split <= (BN == 128 ? 74 : 37)) BN = 64;
  // (128-wide tiles for a layer with half an SM-round of 256-wide ones -- rollout inference, 4096 x 512 = 64 tiles -- were
  // slower: 42.6 vs 35.3 us)
  const uint16_t* Ah = reinterpret_cast<const uint16_t*>(a.A16); const uint16_t* Al = Ah + a.a16_plane;
  const uint16_t* Bh = reinterpret_cast<const uint16_t*>(a.B16); const uint16_t* Bl = Bh + a.b16_plane;
  CUtensorMap tah, tal, tbh, tbl;
  bool ok = p.a_mn ? (make_map_f16(&tah, Ah, a.M, a.K, a.lda, 64, BKh) && make_map_f16(&tal, Al, a.M, a.K, a.lda, 64, BKh))
                   : (make_map_f16(&tah, Ah, a.K, a.M, a.lda, BKh, BM) && make_map_f16(&tal, Al, a.K, a.M, a.lda, BKh, BM));
  ok = ok && (p.b_mn ? (make_map_f16(&tbh, Bh, a.N, a.K, a.ldb, 64, BKh) && make_map_f16(&tbl, Bl, a.N, a.K, a.ldb, 64, BKh))
                     : (make_map_f16(&tbh, Bh, a.K, a.N, a.ldb, BKh, BN) && make_map_f16(&tbl, Bl, a.K, a.N, a.ldb, BKh, BN)));
  if (!ok) { addk_set_error("gemm f16x3: cuTensorMapEncodeTiled rejected an operand"); return ADDK_ERR_ARG; }
  dim3 grid((a.N + BN - 1) / BN, (a.M + BM - 1) / BM, split);
  const int persistent = addk_switches().h3_persistent;
  if (BN == 256 && (persistent || want_planes)) {
    ParamsP pp;
    pp.p = p; pp.a_amax = a.a_amax; pp.b_amax = a.b_amax; pp.comp_per_mma = ph.comp_per_mma; pp.bf16 = 0;
    if (a.C16 && a.c_amax && split == 1 && a.c16_plane > 0 && (a.ldc & 7) == 0 && (reinterpret_cast<uintptr_t>(a.C16) & 15) == 0) {
      pp.p.c_hi = reinterpret_cast<uint16_t*>(a.C16); pp.p.c_plane = a.c16_plane;
    }
    if (a.no_f32) {
      if (!pp.p.c_hi) { addk_set_error("gemm f16x3: no_f32 needs C16 / c16_plane / c_amax (the output lives as fp16 planes)"); return ADDK_ERR_ARG; }
      pp.p.no_f32 = 1;
    }
    if (a.relu_mask_src16) { pp.p.mask16 = reinterpret_cast<const uint16_t*>(a.relu_mask_src16); pp.p.mask = nullptr; }
    { const int rcb = apply_relu_bits(a, pp.p, split); if (rcb != ADDK_OK) return rcb; }
    { const int rcc = apply_colsum_partials(a, pp.p, split); if (rcc != ADDK_OK) return rcc; }
    pp.chunk_kb = addk_switches().h3_chunk_kb;
    g_addk_last_gemm_kernel = ADDK_K_H3_PERSISTENT;
    if (addk_switches().h3_pair && a.M > BM) {      // CTA pairs (cta_group::2): each CTA stages a 128-row half of B
      if (!p.b_mn) ok = make_map_f16(&tbh, Bh, a.K, a.N, a.ldb, BKh, BN / 2) && make_map_f16(&tbl, Bl, a.K, a.N, a.ldb, BKh, BN / 2);
      if (ok) { g_addk_last_gemm_kernel = ADDK_K_H3_PAIR; return launch_h3p<256, false, true>(st, tah, tal, tbh, tbl, pp, a.M, a.N, split); }
    }
    return launch_h3p<256>(st, tah, tal, tbh, tbl, pp, a.M, a.N, split);
  }
  if (a.no_f32 || a.relu_mask_src16 || a.relu_bits_in || a.relu_bits_out || a.colsum_partials) { addk_set_error("gemm f16x3: no_f32 / relu_mask_src16 / relu_bits_* / colsum_partials need the persistent kernel"); return ADDK_ERR_ARG; }
  // (a 256 x 128 instance of the CTA-pair kernel for the 65 .. 128-column layers -- gx = u1.W0, the 1024 x 114 weight
  //  gradients -- measured 27 vs 31 and 43 vs 53 us in isolation and nothing in the whole optimizer step: not kept)
  g_addk_last_gemm_kernel = ADDK_K_H3_TILE;
  if (BN == 256) return launch_h3<256>(st, tah, tal, tbh, tbl, ph, grid);
  if (BN == 128) return launch_h3<128>(st, tah, tal, tbh, tbl, ph, grid);
  return launch_h3<64>(st, tah, tal, tbh, tbl, ph, grid);
}

extern "C" int addk_gemm_is_persistent(const addk_gemm_args* a, int precision) {
  if (!a) return 0;
  const int split = a->split_k > 1 ? a->split_k : 1;
  if (precision == 3) return (a->N > 128 && split == 1 && addk_switches().bf16_persistent) ? 1 : 0;
  if (precision == 4) {
    if (a->N <= 128 || !addk_switches().h3_persistent) return 0;
    return ((long long)((a->M + addk_tc::BM - 1) / addk_tc::BM) * ((a->N + 255) / 256) * split > 37) ? 1 : 0;
  }
  return 0;
}

// Without the legacy tf32x3 kernels (default build) a call the fp16 / bf16 tiles cannot take runs on the exact-fp32
// CUDA-core kernel, which is at least as accurate.
static int gemm_fallback(cudaStream_t st, const addk_gemm_args& a);

int addk_gemm_tc(cudaStream_t st, const addk_gemm_args& a, int precision) {
  if (a.colsum_partials && !((precision == 4 || precision == 3) && addk_gemm_is_persistent(&a, precision))) {
    addk_set_error("gemm: colsum_partials are produced by the persistent f16x3 / bf16 kernels only (addk_gemm_is_persistent)");
    return ADDK_ERR_ARG;
  }
  if ((a.relu_bits_in || a.relu_bits_out) && !((precision == 4 || precision == 3) && addk_gemm_is_persistent(&a, precision))) {
    addk_set_error("gemm: relu_bits_in / relu_bits_out are honoured by the persistent f16x3 / bf16 kernels only (addk_gemm_is_persistent)");
    return ADDK_ERR_ARG;
  }
  if (precision == 4) {
    // fp16 hi/lo planes when the call carries twins that TMA can address (16-byte row pitch = 8 elements, 16-byte
    // aligned planes); otherwise the fallback on the fp32 operands
    const bool ok16 = a.A16 && a.B16 && a.a_amax && a.b_amax &&
                      ((reinterpret_cast<uintptr_t>(a.A16) & 15) == 0) && ((reinterpret_cast<uintptr_t>(a.B16) & 15) == 0) &&
                      ((a.a16_plane & 7) == 0) && ((a.b16_plane & 7) == 0) && a.a16_plane > 0 && a.b16_plane > 0 &&
                      addk_gemm_h3_usable(a);
    if (ok16) return gemm_h3(st, a);
    return gemm_fallback(st, a);
  }
  if (precision == 3) {
    // bf16 tensor-core tiles when the call carries bf16 twins that TMA can address (16-byte row pitch = 8 elements);
    // otherwise the fallback on the fp32 operands, followed by the bf16 copy of the output the caller asked for
    const bool ok16 = a.A16 && a.B16 && ((a.lda & 7) == 0) && ((a.ldb & 7) == 0) && ((reinterpret_cast<uintptr_t>(a.A16) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(a.B16) & 15) == 0) && !a.a_mean && !a.accumulate && a.M >= 16 && a.N >= 16 &&
                      a.K >= 16 && (a.split_k <= 1 || !(a.bias || a.relu || a.relu_mask_src)) && addk_tc::resolve_encode();
    if (ok16) {
      const int rc = gemm_bf16(st, a);
      if (rc >= 0) return rc;
    }
    if (!a.C) { addk_set_error("gemm: bf16 call without an fp32 output cannot fall back"); return ADDK_ERR_ARG; }
    const int rc = gemm_fallback(st, a);
    if (rc != ADDK_OK) return rc;
    if (a.C16 && a.split_k <= 1) {
      const long long n = (long long)a.M * a.N;
      f32_to_bf16_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a.C, reinterpret_cast<uint16_t*>(a.C16), a.M, a.N, a.ldc);
    }
    return ADDK_OK;
  }
#ifndef ADDK_LEGACY_KERNELS
  addk_set_error(precision == 1 || precision == 2
                     ? "gemm: precision tf32x3 / tf32 needs a library built with make LEGACY=1 (superseded by f16x3 / bf16)"
                     : "gemm: unknown precision mode");
  return ADDK_ERR_UNSUPPORTED;
#else
  using namespace addk_tc;
  if (precision != 1 && precision != 2) {
    addk_set_error("gemm: unknown precision mode");
    return ADDK_ERR_UNSUPPORTED;
  }
  const bool aligned = ((a.lda & 3) == 0) && ((a.ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.A) & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(a.B) & 15) == 0);
  g_addk_last_gemm_kernel = ADDK_K_SGEMM;
  if (!aligned || a.a_mean || a.M < 16 || a.N < 16 || a.K < 16 || !resolve_encode()) return addk::sgemm_launch(st, a);
  int split = a.split_k > 1 ? a.split_k : 1;
  if (split > 1 && (a.bias || a.relu || a.relu_mask_src || a.accumulate)) return ADDK_ERR_ARG;
  const bool x3 = precision == 1;
  const int BK = x3 ? 16 : 32;
  const int kb_total = (a.K + BK - 1) / BK;
  int kb_per = (kb_total + split - 1) / split;
  // every slab must own at least one k-block: idle slabs would leave stale data behind
  if ((long long)kb_per * (split - 1) >= kb_total) return addk::sgemm_launch(st, a);
  Params p;
  p.C = a.C; p.ldc = a.ldc; p.M = a.M; p.N = a.N; p.K = a.K; p.bias = a.bias; p.mask = a.relu_mask_src;
  p.ld_mask = a.ld_mask; p.relu = a.relu; p.accumulate = a.accumulate; p.kb_per_split = kb_per;
  p.slab_stride = a.slab_stride > 0 ? a.slab_stride : (long long)a.M * a.ldc;
  p.pair_flags = addk_switches().tc_pair_flags;
  p.dbg = g_addk_stamps;
  p.C16 = nullptr; p.c_amax = nullptr; p.c_hi = nullptr; p.c_plane = 0; p.c_scale = 1.f; p.c16_in_staged = 0; p.no_f32 = 0; p.mask16 = nullptr;
  p.a_mn = a.trans_a ? 1 : 0;          // A given as [K,M]: rows are the contraction index
  p.b_mn = a.trans_b ? 0 : 1;          // B given as [K,N]
  const int BN = a.N > 128 ? 256 : (a.N > 64 ? 128 : 64);
  CUtensorMap ta, tb;
  bool ok = p.a_mn ? make_map(&ta, a.A, a.M, a.K, a.lda, 32, BK, true) : make_map(&ta, a.A, a.K, a.M, a.lda, BK, BM, false);
  ok = ok && (p.b_mn ? make_map(&tb, a.B, a.N, a.K, a.ldb, 32, BK, true) : make_map(&tb, a.B, a.K, a.N, a.ldb, BK, BN, false));
  if (!ok) return addk::sgemm_launch(st, a);
  dim3 grid((a.N + BN - 1) / BN, (a.M + BM - 1) / BM, split);
  if (x3 && BN == 256 && a.M > BM && addk_switches().tc_pair) {
    // CTA-pair kernel: each CTA of the pair stages 128 rows of B -> its tensor-map box has 128 rows
    if (!p.b_mn && !make_map(&tb, a.B, a.K, a.N, a.ldb, BK, Cfg2::BN / 2, false)) return addk::sgemm_launch(st, a);
    g_addk_last_gemm_kernel = ADDK_K_TF32X3_PAIR;
    return launch_x3_pair(st, ta, tb, p, a.M, a.N, split);
  }
  g_addk_last_gemm_kernel = x3 ? ADDK_K_TF32X3 : ADDK_K_TF32;
  if (BN == 256) return x3 ? launch_x3<256>(st, ta, tb, p, grid) : launch<256, false>(st, ta, tb, p, grid);
  if (BN == 128) return x3 ? launch_x3<128>(st, ta, tb, p, grid) : launch<128, false>(st, ta, tb, p, grid);
  return x3 ? launch_x3<64>(st, ta, tb, p, grid) : launch<64, false>(st, ta, tb, p, grid);
#endif
}

static int gemm_fallback(cudaStream_t st, const addk_gemm_args& a) {
#ifdef ADDK_LEGACY_KERNELS
  return addk_gemm_tc(st, a, 1);
#else
  g_addk_last_gemm_kernel = ADDK_K_SGEMM;
  return addk::sgemm_launch(st, a);
#endif
}
