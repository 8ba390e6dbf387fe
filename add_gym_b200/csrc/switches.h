// Experiment / A-B switches of libaddk.so, read from the environment ONCE (first use) -- never on a launch path.
// Defaults are the shipped configuration; every switch is documented where it is consumed.
#pragma once

struct AddkSwitches {
  int tc_pair;            // ADDK_TC_PAIR         (legacy tf32x3) CTA-pair kernel on / off                      default 1
  int tc_pair_flags;      // ADDK_TC_PAIR_FLAGS   (legacy tf32x3) bit0 cluster-scope waits, bit1 relaxed arrives default 2
  int h3_persistent;      // ADDK_H3_PERSISTENT   persistent f16x3 kernel for 256-wide layers                    default 1
  int h3_pair;            // ADDK_H3_PAIR         cta_group::2 CTA pairs for the persistent f16x3 / bf16 kernel          default 1
  int h3_chunk_kb;        // ADDK_H3_CHUNK_KB     k-blocks per accumulator drain                                 default 8
  float h3_comp;          // ADDK_H3_COMP         expected accumulator truncation loss per MMA                   1.7e-8
  int bf16_persistent;    // ADDK_BF16_PERSISTENT persistent one-plane kernel in bf16 mode                       default 1
  int bf16_drop_f32;      // ADDK_BF16_DROP_F32   bf16 mode: layer outputs read only by dense layers / masks / column sums are 16-bit only   default 1
  int h3_planes_only;     // ADDK_H3_PLANES_ONLY  f16x3: such outputs exist only as their fp16 planes (sticky scale + repair launch)     default 1
  int h3_amax_hooks;      // ADDK_H3_AMAX_HOOKS   elementwise producers leave max|x| behind                      default 1
  int h3_fused_planes;    // ADDK_H3_FUSED_PLANES dense-layer epilogue writes the fp16 planes of its output      default 0
  int h3_colpart;         // ADDK_H3_COLPART      split pass leaves bias-gradient column sums behind             default 1
  int h3_relu_bits;       // ADDK_H3_RELU_BITS    optimizer step: ReLU masks travel as bit planes (arena_bits)  default 1
  int fused_tail;         // ADDK_FUSED_TAIL      optimizer step: slab reduction + AdamW + diagnostics row in one launch      default 1
};
const AddkSwitches& addk_switches();
