// f16x3 twins: the power-of-two scale of a tensor and the sticky-scale rule, shared by the dense-layer kernels
// (gemm_tc.cu) and the elementwise kernels that read fp16 planes directly (mlp.cu).
#pragma once
#include <stdint.h>
namespace addk_tc {
// the power-of-two scale of a tensor whose max|x| has the bit pattern `amax_bits`: s = 2^(14 - e), returns s and 1/s
__device__ __forceinline__ void h3_scale(uint32_t amax_bits, float& s, float& inv_s) {
  int E = (int)((amax_bits >> 23) & 0xFFu);
  E = E < 16 ? 16 : (E > 250 ? 250 : E);                       // zero / denormal / huge: any finite scale will do
  s = __uint_as_float((uint32_t)(268 - E) << 23);              // 2^(14 - (E - 127))
  inv_s = __uint_as_float((uint32_t)(E - 14) << 23);
}

// A twin's slot is two words {W, max}: max = bit pattern of max|x|, W = a sticky scale word.  The scale in force is that
// of W while max|x| * s(W) stays inside [2^9, 2^15) -- so a dense layer can write the planes of its OUTPUT in its
// epilogue with the scale its previous output had, before max|x| is known -- and that of 4 * max|x| otherwise (the
// planes are then rewritten by h3_repair_kernel).  Every reader derives the scale from the two words the same way.
__device__ __forceinline__ uint32_t h3_eff_word(uint32_t W, uint32_t amax) {
  const int Ew = (int)((W >> 23) & 0xFFu), Ea = (int)((amax >> 23) & 0xFFu);
  const int d = Ea - Ew + 14;                                  // floor(log2(max|x| * s(W)))
  if (W != 0u && Ew >= 16 && Ew <= 250 && d >= 9 && d <= 14) return W;
  const int En = Ea + 2 > 250 ? 250 : Ea + 2;
  return (amax & 0x007FFFFFu) | ((uint32_t)En << 23);
}
__device__ __forceinline__ void h3_slot_scale(const uint32_t* slot, float& s, float& inv_s) {
  h3_scale(slot ? h3_eff_word(slot[0], slot[1]) : 0x3F800000u, s, inv_s);
}

}  // namespace addk_tc
