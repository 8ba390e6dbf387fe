// Superseded tensor-core kernels, compiled only with -DADDK_LEGACY_KERNELS (make LEGACY=1): single-pass tf32
// (precision "tf32"), the in-kernel hi/lo split tf32x3 kernels (1-CTA and cta_group::2 pair, precision "tf32x3").
// The default library runs f16x3 / bf16 on the kernels of gemm_tc.cu and falls back to the exact-fp32 CUDA-core kernel.
// Included from gemm_tc.cu inside namespace addk_tc.
constexpr int UMMA_K = 8;      // kind::tf32
constexpr float X3_TRUNC_LOSS_PER_MMA = 1.7e-8f;      // measured: the tf32 accumulator truncates toward zero after every instruction

// One stage of the ring.  A: 128 (rows | columns) x BK k;  B: BN x BK k.  tf32x3 adds the "lo" halves and uses
// BK = 16 (64-byte rows) so that four stages still fit; the single-pass mode uses BK = 32 (128-byte rows).
template <int BN, bool X3>
struct Cfg {
  static constexpr int BK = X3 ? 16 : 32;
  static constexpr int A_BYTES = BM * BK * 4;
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = (A_BYTES + B_BYTES) * (X3 ? 2 : 1);
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 6 ? 6 : (200 * 1024) / STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
  // tf32x3 keeps the small cross terms (lo.hi + hi.lo) in a second accumulator: the tensor core truncates its
  // fp32 accumulator after every instruction, so three accumulations per k-step into ONE accumulator would
  // triple that bias; the cross-term accumulator is 2^-11 smaller and its truncation is negligible.
  static constexpr int TMEM_COLS = (X3 ? 2 : 1) * (BN < 32 ? 32 : BN);
  // K-major tiles: 128-byte rows -> SWIZZLE_128B (UMMA layout 2), 64-byte rows -> SWIZZLE_64B (layout 4);
  // 8-row groups are 8 * row bytes apart (SBO).  MN-major tiles: one TMA box = 32 MN x BK k (BK * 128 bytes),
  // 128B swizzle with 32-byte atoms (layout 1), MN atoms one box apart (LBO), 4-k groups 512 B apart (SBO).
  static constexpr uint32_t K_LAYOUT = BK == 32 ? 2u : 4u;
  static constexpr uint32_t K_SBO = 8u * BK * 4u;
  static constexpr uint32_t MN_BOX_BYTES = BK * 128u;
};

template <int BN, bool X3>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
  using C = Cfg<BN, X3>;
  constexpr int BK = C::BK;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* const base_ptr = smem_raw + (base - raw_u32);
  const uint32_t bars = base + C::STAGES * C::STAGE_BYTES;      // full[S] | empty[S] | ready[S] | tmem_full | tmem_ptr
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
  auto ready_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + s); };
  const uint32_t tmem_full_bar = bars + 8u * (3 * C::STAGES);
  const uint32_t tmem_ptr_addr = bars + 8u * (3 * C::STAGES + 1);
  auto a_hi = [&](int s) { return base + (uint32_t)s * C::STAGE_BYTES; };
  auto b_hi = [&](int s) { return a_hi(s) + C::A_BYTES; };
  auto a_lo = [&](int s) { return b_hi(s) + C::B_BYTES; };
  auto b_lo = [&](int s) { return a_lo(s) + C::A_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kb_total = (p.K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(kb_total, kb_begin + p.kb_per_split);
  const int num_kb = kb_end - kb_begin;          // host guarantees >= 1

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(ready_bar(s), 128);
    }
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
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % C::STAGES;
        const uint32_t ph = (uint32_t)(i / C::STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), C::A_BYTES + C::B_BYTES);
        const int k0 = (kb_begin + i) * BK;
        if (!p.a_mn) {
          tma_load_2d(a_hi(s), &tmA, full_bar(s), k0, m0);                      // box {BK k, 128 rows}
        } else {
#pragma unroll
          for (int j = 0; j < BM / 32; ++j)                                      // box {32 m, BK k}
            tma_load_2d(a_hi(s) + j * C::MN_BOX_BYTES, &tmA, full_bar(s), m0 + 32 * j, k0);
        }
        if (!p.b_mn) {
          tma_load_2d(b_hi(s), &tmB, full_bar(s), k0, n0);                      // box {BK k, BN rows}
        } else {
#pragma unroll
          for (int j = 0; j < BN / 32; ++j)
            tma_load_2d(b_hi(s) + j * C::MN_BOX_BYTES, &tmB, full_bar(s), n0 + 32 * j, k0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D fp32, A/B tf32, majors, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.a_mn ? 1 : 0) << 15) |
                             ((uint32_t)(p.b_mn ? 1 : 0) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t a_lbo = p.a_mn ? C::MN_BOX_BYTES : 16u, b_lbo = p.b_mn ? C::MN_BOX_BYTES : 16u;
      const uint32_t a_sbo = p.a_mn ? 512u : C::K_SBO, b_sbo = p.b_mn ? 512u : C::K_SBO;
      const uint32_t a_lay = p.a_mn ? 1u : C::K_LAYOUT, b_lay = p.b_mn ? 1u : C::K_LAYOUT;
      const uint32_t a_kstep = p.a_mn ? 1024u : 32u, b_kstep = p.b_mn ? 1024u : 32u;   // 8 k per MMA
      uint32_t acc = 0, acc_x = 0;
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % C::STAGES;
        const uint32_t ph = (uint32_t)(i / C::STAGES) & 1u;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        // main term: the tensor core truncates the fp32 operands to tf32 itself, so hi(x) is the landed tile as is
#pragma unroll
        for (int ks = 0; ks < BK / UMMA_K; ++ks) {
          umma_tf32(tmem_base, smem_desc(a_hi(s) + ks * a_kstep, a_lbo, a_sbo, a_lay),
                    smem_desc(b_hi(s) + ks * b_kstep, b_lbo, b_sbo, b_lay), idesc, acc);
          acc = 1;
        }
        if (X3) {
          mbar_wait(ready_bar(s), ph);         // lo tiles written by the splitter warps
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < BK / UMMA_K; ++ks) {
            const uint64_t dah = smem_desc(a_hi(s) + ks * a_kstep, a_lbo, a_sbo, a_lay);
            const uint64_t dbh = smem_desc(b_hi(s) + ks * b_kstep, b_lbo, b_sbo, b_lay);
            const uint64_t dal = smem_desc(a_lo(s) + ks * a_kstep, a_lbo, a_sbo, a_lay);
            const uint64_t dbl = smem_desc(b_lo(s) + ks * b_kstep, b_lbo, b_sbo, b_lay);
            umma_tf32(tmem_base + BN, dal, dbh, idesc, acc_x);
            acc_x = 1;
            umma_tf32(tmem_base + BN, dah, dbl, idesc, acc_x);
          }
        }
        umma_commit(empty_bar(s));          // stage reusable once these MMAs have read it
      }
      umma_commit(tmem_full_bar);           // accumulators complete
    }
  } else {
    // ===================== splitter (tf32x3) + epilogue: warps 2..5 =====================
    const int t = threadIdx.x - 64;          // 0..127
    if (X3) {
      // lo = x - tf32_trunc(x), elementwise (so the swizzled placement does not matter; A and B are contiguous
      // in the stage and so are their lo twins).  The hi tile is left untouched: the MMA reads it concurrently.
      constexpr int N4 = (C::A_BYTES + C::B_BYTES) / 16;
      constexpr int PER = N4 / 128;
      static_assert(N4 % 128 == 0, "tile size");
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % C::STAGES;
        const uint32_t ph = (uint32_t)(i / C::STAGES) & 1u;
        mbar_wait(full_bar(s), ph);
        const float4* src = reinterpret_cast<const float4*>(base_ptr + (size_t)s * C::STAGE_BYTES);
        float4* dst = reinterpret_cast<float4*>(base_ptr + (size_t)s * C::STAGE_BYTES + C::A_BYTES + C::B_BYTES);
#pragma unroll
        for (int j0 = 0; j0 < PER; j0 += 4) {
          float4 x[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) if (j0 + u < PER) x[u] = src[t + 128 * (j0 + u)];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (j0 + u < PER) {
              float4 l;
              l.x = x[u].x - __uint_as_float(__float_as_uint(x[u].x) & 0xFFFFE000u);
              l.y = x[u].y - __uint_as_float(__float_as_uint(x[u].y) & 0xFFFFE000u);
              l.z = x[u].z - __uint_as_float(__float_as_uint(x[u].z) & 0xFFFFE000u);
              l.w = x[u].w - __uint_as_float(__float_as_uint(x[u].w) & 0xFFFFE000u);
              dst[t + 128 * (j0 + u)] = l;
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        mbar_arrive(ready_bar(s));
      }
    }
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int row = m0 + 32 * q + lane;
    float* Cz = p.C + (size_t)blockIdx.z * p.slab_stride;
    const bool vec = epilogue_vec_ok(p, Cz);
    // Each warp owns rows [32q, 32q+32) of the tile: blocks of up to 128 columns go TMEM -> registers -> a 16 KB
    // staging tile per warp (the operand stages are idle by now) -> full-row stores (store_staged).
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    constexpr int CWB = BN < 128 ? BN : 128;
    float4* stg = reinterpret_cast<float4*>(base_ptr + (32 * CWB * 4) * q);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += CWB) {
      if (n0 + c0 >= p.N) break;             // warp-uniform
#pragma unroll
      for (int cc = 0; cc < CWB / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(c0 + cc * 32), v);
        if (X3) {
          uint32_t w[32];
          tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(BN + c0 + cc * 32), w);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
        }
        if (vec) {
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4)
            stage_put<CWB>(stg, lane, cc * 8 + c4, __uint_as_float(v[4 * c4]), __uint_as_float(v[4 * c4 + 1]),
                           __uint_as_float(v[4 * c4 + 2]), __uint_as_float(v[4 * c4 + 3]));
        } else if (row < p.M) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          store_row_scalar(p, Cz, row, n0 + c0 + cc * 32, f);
        }
      }
      if (vec) {
        __syncwarp();
        store_staged<CWB>(p, Cz, stg, lane, m0 + 32 * q, n0 + c0);
        __syncwarp();
      }
    }
  }
  // ===================== teardown =====================
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
  }
}


// 2-D fp32 tensor map: memory [outer, inner] with `ld` floats between rows; box {32, box_rows}, 128-byte swizzle.
static bool make_map(CUtensorMap* map, const float* ptr, long long inner, long long outer, long long ld, int box_inner,
                     int box_rows, bool mn_major) {
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                 : (box_inner == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B),
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}


// ---------------------------------------------------------------------------------------------------------------
// tf32x3 ("fp32-parity") kernel.  Same TMA / tcgen05 pipeline as above plus two things the accuracy bar needs:
//  * the cross terms lo.hi + hi.lo go to a second TMEM accumulator;
//  * the main accumulator is DRAINED into fp32 registers every X3_CHUNK_KB k-blocks (K = 256): the tensor core
//    truncates its accumulator after every instruction (measured bias -1.64e-8 per accumulated MMA, i.e. -2.1e-6 at
//    K = 1024), so the tensor core only ever sums 32 instructions and the CUDA cores add the chunks with
//    round-to-nearest -> 5e-7, the level of an fp32 FMA loop.
// 10 warps: 0 = TMA producer, 1 = MMA issuer / TMEM allocator, 2..9 = workers (hi/lo split of every landed tile,
// chunk drains, epilogue).  Worker w owns TMEM lanes 32*(w%4).. and column half (w-2)/4, BN/2 running sums per thread.
// ---------------------------------------------------------------------------------------------------------------
constexpr int X3_CHUNK_KB = 16;
// Expected truncation loss of the tensor core's accumulator per accumulated instruction, relative to the chunk sum
// (measured on B200 with tf32-exact operands: -1.68e-8 .. -2.1e-8 per instruction for 4..2048 instructions,
// tools/tc_accuracy.py).  The drain adds it back, which removes the systematic part of the bias (-5.4e-7 per
// 32-instruction chunk) and leaves the random part (~3e-7).

template <int BN>
__global__ void __launch_bounds__(X3_THREADS, 1)
gemm_tc_x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
  using C = Cfg<BN, true>;
  constexpr int BK = C::BK;
  constexpr int CPW = BN / 2;                     // accumulator columns per worker thread
  constexpr int NCH = CPW / 32;                   // 32-column chunks per worker
  static_assert(CPW % 32 == 0, "BN must be a multiple of 64");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* const base_ptr = smem_raw + (base - raw_u32);
  const uint32_t bars = base + C::STAGES * C::STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
  auto ready_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + s); };
  const uint32_t tmem_full_bar = bars + 8u * (3 * C::STAGES);
  const uint32_t chunk_full_bar = bars + 8u * (3 * C::STAGES + 1);
  const uint32_t chunk_empty_bar = bars + 8u * (3 * C::STAGES + 2);
  const uint32_t tmem_ptr_addr = bars + 8u * (3 * C::STAGES + 3);
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
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(ready_bar(s), 256);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(chunk_full_bar, 1);
    mbar_init(chunk_empty_bar, 256);
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
        if ((p.pair_flags & 8) && i >= C::STAGES) { mbar_arrive(full_bar(s)); continue; }   // experiment: no TMA traffic
        mbar_expect_tx(full_bar(s), C::A_BYTES + C::B_BYTES);
        const int k0 = (kb_begin + i) * BK;
        if (!p.a_mn) {
          tma_load_2d(a_hi(s), &tmA, full_bar(s), k0, m0);
        } else {
#pragma unroll
          for (int j = 0; j < BM / 32; ++j) tma_load_2d(a_hi(s) + j * C::MN_BOX_BYTES, &tmA, full_bar(s), m0 + 32 * j, k0);
        }
        if (!p.b_mn) {
          tma_load_2d(b_hi(s), &tmB, full_bar(s), k0, n0);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) tma_load_2d(b_hi(s) + j * C::MN_BOX_BYTES, &tmB, full_bar(s), n0 + 32 * j, k0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.a_mn ? 1 : 0) << 15) |
                             ((uint32_t)(p.b_mn ? 1 : 0) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t a_lbo = p.a_mn ? C::MN_BOX_BYTES : 16u, b_lbo = p.b_mn ? C::MN_BOX_BYTES : 16u;
      const uint32_t a_sbo = p.a_mn ? 512u : C::K_SBO, b_sbo = p.b_mn ? 512u : C::K_SBO;
      const uint32_t a_lay = p.a_mn ? 1u : C::K_LAYOUT, b_lay = p.b_mn ? 1u : C::K_LAYOUT;
      const uint32_t a_kstep = p.a_mn ? 1024u : 32u, b_kstep = p.b_mn ? 1024u : 32u;
      uint32_t acc = 0, acc_x = 0;
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % C::STAGES;
        const uint32_t ph = (uint32_t)(i / C::STAGES) & 1u;
        const bool new_chunk = (i % X3_CHUNK_KB == 0) && i > 0;
        // cross terms first at a chunk boundary: they go to the other accumulator and overlap the drain
        mbar_wait(ready_bar(s), ph);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < BK / UMMA_K; ++ks) {
          const uint64_t dah = smem_desc(a_hi(s) + ks * a_kstep, a_lbo, a_sbo, a_lay);
          const uint64_t dbh = smem_desc(b_hi(s) + ks * b_kstep, b_lbo, b_sbo, b_lay);
          const uint64_t dal = smem_desc(a_lo(s) + ks * a_kstep, a_lbo, a_sbo, a_lay);
          const uint64_t dbl = smem_desc(b_lo(s) + ks * b_kstep, b_lbo, b_sbo, b_lay);
          umma_tf32(tmem_base + BN, dal, dbh, idesc, acc_x);
          acc_x = 1;
          umma_tf32(tmem_base + BN, dah, dbl, idesc, acc_x);
        }
        if (new_chunk && !(p.pair_flags & 4)) {   // the workers have copied the previous chunk out of the main accumulator
          mbar_wait(chunk_empty_bar, (uint32_t)(i / X3_CHUNK_KB - 1) & 1u);
          tc_fence_after();
          acc = 0;
        }
#pragma unroll
        for (int ks = 0; ks < BK / UMMA_K; ++ks) {
          umma_tf32(tmem_base, smem_desc(a_hi(s) + ks * a_kstep, a_lbo, a_sbo, a_lay),
                    smem_desc(b_hi(s) + ks * b_kstep, b_lbo, b_sbo, b_lay), idesc, acc);
          acc = 1;
        }
        umma_commit(empty_bar(s));
        if (((i + 1) % X3_CHUNK_KB == 0) && (i + 1 < num_kb)) umma_commit(chunk_full_bar);
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ===================== workers: warps 2..9 =====================
    const int t = threadIdx.x - 64;          // 0..255
    const int q = warp & 3;                  // TMEM lane quarter
    const int half = (warp - 2) >> 2;        // column half
    const uint32_t t_main = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(half * CPW);
    float acc[CPW];
#pragma unroll
    for (int j = 0; j < CPW; ++j) acc[j] = 0.f;
    constexpr int N4 = (C::A_BYTES + C::B_BYTES) / 16;
    constexpr int PER = N4 / 256;
    static_assert(N4 % 256 == 0, "tile size");
    for (int i = 0; i < num_kb; ++i) {
      const int s = i % C::STAGES;
      const uint32_t ph = (uint32_t)(i / C::STAGES) & 1u;
      mbar_wait(full_bar(s), ph);
      const float4* src = reinterpret_cast<const float4*>(base_ptr + (size_t)s * C::STAGE_BYTES);
      float4* dst = reinterpret_cast<float4*>(base_ptr + (size_t)s * C::STAGE_BYTES + C::A_BYTES + C::B_BYTES);
      if (p.pair_flags & 4) { mbar_arrive(ready_bar(s)); continue; }   // experiment: no split work (and no drain)
      float4 x[PER];
#pragma unroll
      for (int u = 0; u < PER; ++u) x[u] = src[t + 256 * u];
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        float4 l;
        l.x = x[u].x - __uint_as_float(__float_as_uint(x[u].x) & 0xFFFFE000u);
        l.y = x[u].y - __uint_as_float(__float_as_uint(x[u].y) & 0xFFFFE000u);
        l.z = x[u].z - __uint_as_float(__float_as_uint(x[u].z) & 0xFFFFE000u);
        l.w = x[u].w - __uint_as_float(__float_as_uint(x[u].w) & 0xFFFFE000u);
        dst[t + 256 * u] = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(ready_bar(s));
      if ((i % X3_CHUNK_KB == 0) && i > 0) {   // drain the chunk that ended with k-block i-1
        mbar_wait(chunk_full_bar, (uint32_t)(i / X3_CHUNK_KB - 1) & 1u);
        tc_fence_after();
        const float comp = X3_TRUNC_LOSS_PER_MMA * (float)(X3_CHUNK_KB * (BK / UMMA_K));
#pragma unroll
        for (int cc = 0; cc < NCH; ++cc) {
          uint32_t v[32];
          tmem_ld32(t_main + (uint32_t)(cc * 32), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[cc * 32 + j] += fmaf(__uint_as_float(v[j]), comp, __uint_as_float(v[j]));
        }
        tc_fence_before();
        mbar_arrive(chunk_empty_bar);
      }
    }
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const float comp_last = X3_TRUNC_LOSS_PER_MMA * (float)((((num_kb - 1) % X3_CHUNK_KB) + 1) * (BK / UMMA_K));
#pragma unroll
    for (int cc = 0; cc < NCH; ++cc) {         // last chunk of the main accumulator + the cross-term accumulator
      uint32_t v[32];
      tmem_ld32(t_main + (uint32_t)(cc * 32), v);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[cc * 32 + j] += fmaf(__uint_as_float(v[j]), comp_last, __uint_as_float(v[j]));
      tmem_ld32(t_main + (uint32_t)(BN + cc * 32), v);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[cc * 32 + j] += __uint_as_float(v[j]);
    }
    // ---- epilogue: this warp's 32 x CPW accumulators -> staging tile -> full-row stores (store_staged)
    float* Cz = p.C + (size_t)blockIdx.z * p.slab_stride;
    const bool vec = epilogue_vec_ok(p, Cz);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    float4* stg = reinterpret_cast<float4*>(base_ptr + (32 * CPW * 4) * (warp - 2));
    const int row = m0 + 32 * q + lane;
    const int cw0 = n0 + half * CPW;
    if (cw0 < p.N) {                           // warp-uniform
      if (vec) {
#pragma unroll
        for (int sl = 0; sl < CPW / 4; ++sl) stage_put<CPW>(stg, lane, sl, acc[4 * sl], acc[4 * sl + 1], acc[4 * sl + 2], acc[4 * sl + 3]);
        __syncwarp();
        store_staged<CPW>(p, Cz, stg, lane, m0 + 32 * q, cw0);
      } else if (row < p.M) {
#pragma unroll
        for (int cc = 0; cc < NCH; ++cc) store_row_scalar(p, Cz, row, cw0 + cc * 32, acc + cc * 32);
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

template <int BN>
static int launch_x3(cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, dim3 grid) {
  using C = Cfg<BN, true>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_tc_x3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess) {
      addk_set_error("gemm_tc: cannot raise the dynamic shared memory limit");
      return ADDK_ERR_LAUNCH;
    }
    configured = true;
  }
  gemm_tc_x3_kernel<BN><<<grid, X3_THREADS, C::SMEM_BYTES, st>>>(ta, tb, p);
  return ADDK_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// tf32x3, CTA-pair version (tcgen05 cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 tile.
// Each CTA stages its own 128 rows of A and its own 128-row half of B; the leader CTA's single thread issues
// M = 256 MMAs that read both CTAs' shared memory, so per CTA the operand traffic per k-block drops from
// (128 + 256) to (128 + 128) rows -- less L2->smem traffic, less splitting work, 1/3 fewer operand bytes per MMA.
// Everything else (hi/lo split, second accumulator for the cross terms, chunked drain, coalesced epilogue) is the
// 1-CTA kernel above; each CTA drains / stores its own 128 accumulator rows.
// Barriers: full/empty are CTA-local (local TMA; multicast tcgen05.commit frees the stage in both CTAs);
// ready / chunk_empty live in the leader and collect one arrival per worker warp of BOTH CTAs (remote arrive).
// ---------------------------------------------------------------------------------------------------------------
struct Cfg2 {
  static constexpr int BN = 256;                  // tile N; each CTA stages BN/2 rows of B
  static constexpr int BK = 16;
  static constexpr int A_BYTES = BM * BK * 4;     // 8 KB
  static constexpr int B_BYTES = (BN / 2) * BK * 4;
  static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);
  static constexpr int STAGES = 6;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 512;
  static constexpr uint32_t K_LAYOUT = 4u;
  static constexpr uint32_t K_SBO = 8u * BK * 4u;
  static constexpr uint32_t MN_BOX_BYTES = BK * 128u;
};

__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void umma_tf32_2cta(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(X3_THREADS, 1)
gemm_tc_x3_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
  using C = Cfg2;
  constexpr int BK = C::BK, BN = C::BN;
  constexpr int CPW = BN / 2, NCH = CPW / 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* const base_ptr = smem_raw + (base - raw_u32);
  const uint32_t bars = base + C::STAGES * C::STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
  auto ready_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + s); };
  const uint32_t tmem_full_bar = bars + 8u * (3 * C::STAGES);
  const uint32_t chunk_full_bar = bars + 8u * (3 * C::STAGES + 1);
  const uint32_t chunk_empty_bar = bars + 8u * (3 * C::STAGES + 2);
  const uint32_t tmem_ptr_addr = bars + 8u * (3 * C::STAGES + 3);
  auto a_hi = [&](int s) { return base + (uint32_t)s * C::STAGE_BYTES; };
  auto b_hi = [&](int s) { return a_hi(s) + C::A_BYTES; };
  auto a_lo = [&](int s) { return b_hi(s) + C::B_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int m0 = (blockIdx.y * 2 + (int)rank) * BM;          // this CTA's 128 accumulator rows
  const int n0 = (blockIdx.x >> 1) * BN;                      // tile columns
  const int nb0 = n0 + (int)rank * (BN / 2);                  // this CTA's half of B
  const int kb_total = (p.K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(kb_total, kb_begin + p.kb_per_split);
  const int num_kb = kb_end - kb_begin;
  long long* const dbg = (p.dbg && leader && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? p.dbg : nullptr;
  if (dbg && threadIdx.x == 32) dbg[0] = clock64();

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(ready_bar(s), 16);            // 8 worker warps x 2 CTAs (only the leader's copy is used)
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(chunk_full_bar, 1);
    mbar_init(chunk_empty_bar, 16);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"((uint32_t)C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr) : "memory");
  if (dbg && threadIdx.x == 32) dbg[1] = clock64();

  if (warp == 0) {
    // ---- TMA producer (each CTA loads its own tiles)
    if (lane == 0) {
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % C::STAGES;
        const uint32_t ph = (uint32_t)(i / C::STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        if ((p.pair_flags & 8) && i >= C::STAGES) { mbar_arrive(full_bar(s)); continue; }   // experiment: no TMA traffic
        mbar_expect_tx(full_bar(s), C::A_BYTES + C::B_BYTES);
        const int k0 = (kb_begin + i) * BK;
        if (!p.a_mn) {
          tma_load_2d(a_hi(s), &tmA, full_bar(s), k0, m0);
        } else {
#pragma unroll
          for (int j = 0; j < BM / 32; ++j) tma_load_2d(a_hi(s) + j * C::MN_BOX_BYTES, &tmA, full_bar(s), m0 + 32 * j, k0);
        }
        if (!p.b_mn) {
          tma_load_2d(b_hi(s), &tmB, full_bar(s), k0, nb0);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(b_hi(s) + j * C::MN_BOX_BYTES, &tmB, full_bar(s), nb0 + 32 * j, k0);
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: leader CTA only.  The WHOLE warp runs the loop so that the descriptors stay warp-uniform
    // (uniform registers, no per-MMA R2UR / address arithmetic); one elected lane issues.  Measured: with a single
    // divergent thread computing descriptors the issue thread, not the tensor pipe, paced the kernel (230 cycles per
    // 128x256x8 MMA against a floor of 128).
    if (leader) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.a_mn ? 1 : 0) << 15) |
                             ((uint32_t)(p.b_mn ? 1 : 0) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
      const uint32_t a_lbo = p.a_mn ? C::MN_BOX_BYTES : 16u, b_lbo = p.b_mn ? C::MN_BOX_BYTES : 16u;
      const uint32_t a_sbo = p.a_mn ? 512u : C::K_SBO, b_sbo = p.b_mn ? 512u : C::K_SBO;
      const uint32_t a_lay = p.a_mn ? 1u : C::K_LAYOUT, b_lay = p.b_mn ? 1u : C::K_LAYOUT;
      // descriptors of stage 0, k-step 0; everything else is a constant added to the 14-bit start-address field
      const uint64_t dA0 = smem_desc(a_hi(0), a_lbo, a_sbo, a_lay);
      const uint64_t dB0 = smem_desc(b_hi(0), b_lbo, b_sbo, b_lay);
      const uint64_t a_k16 = p.a_mn ? (1024u >> 4) : (32u >> 4), b_k16 = p.b_mn ? (1024u >> 4) : (32u >> 4);
      constexpr uint64_t LO16 = (C::A_BYTES + C::B_BYTES) >> 4, STAGE16 = C::STAGE_BYTES >> 4;
      const bool issuer = elect_one();
      const bool no_drain = (p.pair_flags & 4) != 0;
      const bool pair_wait_cluster = (p.pair_flags & 1) != 0;
      uint32_t acc = 0, acc_x = 0, ph = 0, chunk_par = 0;
      int chunk_left = X3_CHUNK_KB;       // k-blocks left in the current drain chunk
      for (int i = 0; i < num_kb; ph ^= 1u) {
#pragma unroll
        for (int s = 0; s < C::STAGES; ++s) {
          if (i >= num_kb) break;
          if (pair_wait_cluster) mbar_wait_cluster(ready_bar(s), ph); else mbar_wait(ready_bar(s), ph);   // both CTAs: tiles landed, lo halves written
          tc_fence_after();
          if (dbg && i == 0 && lane == 0) dbg[2] = clock64();
          const uint64_t dah = dA0 + s * STAGE16, dbh = dB0 + s * STAGE16;
          if (issuer) {
#pragma unroll
            for (int ks = 0; ks < BK / UMMA_K; ++ks) {
              umma_tf32_2cta(tmem_base + BN, dah + LO16 + ks * a_k16, dbh + ks * b_k16, idesc, acc_x);
              acc_x = 1;
              umma_tf32_2cta(tmem_base + BN, dah + ks * a_k16, dbh + LO16 + ks * b_k16, idesc, acc_x);
            }
          }
          if (chunk_left == 0) {          // the workers have copied the previous chunk out of the main accumulator
            if (!no_drain) {
              if (pair_wait_cluster) mbar_wait_cluster(chunk_empty_bar, chunk_par); else mbar_wait(chunk_empty_bar, chunk_par);
              tc_fence_after();
              acc = 0;
            }
            chunk_par ^= 1u;
            chunk_left = X3_CHUNK_KB;
          }
          --chunk_left;
          ++i;
          if (issuer) {
#pragma unroll
            for (int ks = 0; ks < BK / UMMA_K; ++ks) {
              umma_tf32_2cta(tmem_base, dah + ks * a_k16, dbh + ks * b_k16, idesc, acc);
              acc = 1;
            }
            umma_commit_2cta(empty_bar(s));
            if (chunk_left == 0 && i < num_kb) umma_commit_2cta(chunk_full_bar);
          }
          __syncwarp();
        }
      }
      if (issuer) umma_commit_2cta(tmem_full_bar);
      __syncwarp();
      if (dbg && lane == 0) dbg[3] = clock64();
    }
  } else {
    // ---- workers: warps 2..9
    const int t = threadIdx.x - 64;
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t t_main = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(half * CPW);
    float acc[CPW];
#pragma unroll
    for (int j = 0; j < CPW; ++j) acc[j] = 0.f;
    constexpr int N4 = (C::A_BYTES + C::B_BYTES) / 16;
    constexpr int PER = N4 / 256;
    static_assert(N4 % 256 == 0, "tile size");
    uint32_t ready_remote[C::STAGES];
#pragma unroll
    for (int s = 0; s < C::STAGES; ++s) ready_remote[s] = mapa_rank0(ready_bar(s));
    const uint32_t chunk_empty_remote = mapa_rank0(chunk_empty_bar);
    for (int i = 0; i < num_kb; ++i) {
      const int s = i % C::STAGES;
      const uint32_t ph = (uint32_t)(i / C::STAGES) & 1u;
      mbar_wait(full_bar(s), ph);
      const float4* src = reinterpret_cast<const float4*>(base_ptr + (size_t)s * C::STAGE_BYTES);
      float4* dst = reinterpret_cast<float4*>(base_ptr + (size_t)s * C::STAGE_BYTES + C::A_BYTES + C::B_BYTES);
      if (p.pair_flags & 4) {   // experiment: no split work (and no drain)
        __syncwarp();
        if (lane == 0) {
          uint32_t ra = ready_remote[0];
#pragma unroll
          for (int ss = 1; ss < C::STAGES; ++ss) if (s == ss) ra = ready_remote[ss];
          mbar_arrive_cluster_relaxed(ra);
        }
        continue;
      }
      float4 x[PER];
#pragma unroll
      for (int u = 0; u < PER; ++u) x[u] = src[t + 256 * u];
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        float4 l;
        l.x = x[u].x - __uint_as_float(__float_as_uint(x[u].x) & 0xFFFFE000u);
        l.y = x[u].y - __uint_as_float(__float_as_uint(x[u].y) & 0xFFFFE000u);
        l.z = x[u].z - __uint_as_float(__float_as_uint(x[u].z) & 0xFFFFE000u);
        l.w = x[u].w - __uint_as_float(__float_as_uint(x[u].w) & 0xFFFFE000u);
        dst[t + 256 * u] = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        uint32_t ra = ready_remote[0];
#pragma unroll
        for (int ss = 1; ss < C::STAGES; ++ss) if (s == ss) ra = ready_remote[ss];
        if (p.pair_flags & 2) mbar_arrive_cluster_relaxed(ra); else mbar_arrive_cluster(ra);
      }
      if ((i % X3_CHUNK_KB == 0) && i > 0) {
        mbar_wait(chunk_full_bar, (uint32_t)(i / X3_CHUNK_KB - 1) & 1u);
        tc_fence_after();
        const float comp = X3_TRUNC_LOSS_PER_MMA * (float)(X3_CHUNK_KB * (BK / UMMA_K));
#pragma unroll
        for (int cc = 0; cc < NCH; ++cc) {
          uint32_t v[32];
          tmem_ld32(t_main + (uint32_t)(cc * 32), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[cc * 32 + j] += fmaf(__uint_as_float(v[j]), comp, __uint_as_float(v[j]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (p.pair_flags & 2) mbar_arrive_cluster_relaxed(chunk_empty_remote); else mbar_arrive_cluster(chunk_empty_remote); }
      }
    }
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (dbg && threadIdx.x == 64) dbg[4] = clock64();
    const float comp_last = X3_TRUNC_LOSS_PER_MMA * (float)((((num_kb - 1) % X3_CHUNK_KB) + 1) * (BK / UMMA_K));
#pragma unroll
    for (int cc = 0; cc < NCH; ++cc) {
      uint32_t v[32];
      tmem_ld32(t_main + (uint32_t)(cc * 32), v);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[cc * 32 + j] += fmaf(__uint_as_float(v[j]), comp_last, __uint_as_float(v[j]));
      tmem_ld32(t_main + (uint32_t)(BN + cc * 32), v);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[cc * 32 + j] += __uint_as_float(v[j]);
    }
    // ---- epilogue: this warp's 32 x 128 accumulators -> 16 KB staging tile -> full-row stores (store_staged)
    float* Cz = p.C + (size_t)blockIdx.z * p.slab_stride;
    const bool vec = epilogue_vec_ok(p, Cz);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    float4* stg = reinterpret_cast<float4*>(base_ptr + (32 * CPW * 4) * (warp - 2));
    const int row = m0 + 32 * q + lane;
    const int cw0 = n0 + half * CPW;                 // first column of this warp
    if (cw0 < p.N) {
      if (vec) {
#pragma unroll
        for (int sl = 0; sl < CPW / 4; ++sl) stage_put<CPW>(stg, lane, sl, acc[4 * sl], acc[4 * sl + 1], acc[4 * sl + 2], acc[4 * sl + 3]);
        __syncwarp();
        store_staged<CPW>(p, Cz, stg, lane, m0 + 32 * q, cw0);
      } else if (row < p.M) {
#pragma unroll
        for (int cc = 0; cc < NCH; ++cc) store_row_scalar(p, Cz, row, cw0 + cc * 32, acc + cc * 32);
      }
    }
  }
  if (dbg && threadIdx.x == 64) dbg[5] = clock64();
  tc_fence_before();
  cluster_sync_all();
  if (dbg && threadIdx.x == 32) dbg[6] = clock64();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
  }
}

static int launch_x3_pair(cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, int M, int N, int split) {
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_tc_x3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2::SMEM_BYTES) != cudaSuccess) {
      addk_set_error("gemm_tc: cannot raise the dynamic shared memory limit");
      return ADDK_ERR_LAUNCH;
    }
    configured = true;
  }
  dim3 grid(2 * ((N + Cfg2::BN - 1) / Cfg2::BN), (M + 2 * BM - 1) / (2 * BM), split);
  gemm_tc_x3_pair_kernel<<<grid, X3_THREADS, Cfg2::SMEM_BYTES, st>>>(ta, tb, p);
  return ADDK_OK;
}


template <int BN, bool X3>
static int launch(cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, dim3 grid) {
  using C = Cfg<BN, X3>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_tc_kernel<BN, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess) {
      addk_set_error("gemm_tc: cannot raise the dynamic shared memory limit");
      return ADDK_ERR_LAUNCH;
    }
    configured = true;
  }
  gemm_tc_kernel<BN, X3><<<grid, NTHREADS, C::SMEM_BYTES, st>>>(ta, tb, p);
  return ADDK_OK;
}

