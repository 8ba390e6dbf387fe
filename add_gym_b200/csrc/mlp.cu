// PPO actor-critic + ADD discriminator: rollout inference, batched evaluation and one full optimizer
// step (gather -> forward -> losses -> backward incl. the gradient-penalty double backward -> AdamW),
// issued from native host code as a fixed sequence of ~60 launches per minibatch.
//
// Replaces (reference add_gym/learning/):
//   ExperienceBuffer.sample                       experience_buffer.py:74-113
//   Normalizer.normalize / DiffNormalizer.normalize  normalizer.py:107-110, diff_normalizer.py:56-60
//   PPOModel.eval_actor/eval_critic, ADDModel.eval_disc  ppo_model.py:13-21, add/add_model.py:12-15
//   DistributionGaussianDiag.sample/log_prob       distribution_gaussian_diag.py:84-94
//   PPOAgent._decide_action                        ppo_agent.py:72-104
//   PPOAgent._compute_{critic,actor}_loss, BaseAgent._compute_action_bound_loss
//                                                  ppo_agent.py:209-261, base_agent.py:522-546
//   ADDAgent._compute_disc_loss, AMPAgent._disc_loss_pos/_neg/_compute_disc_acc
//                                                  add/add_agent.py:141-202, amp_agent.py:177-192
//   MPOptimizer.step (zero_grad, backward, AdamW)  mp_optimizer.py:14-23
//
// Backward is written out by hand.  For the discriminator D(x) = w3.relu(W2 relu(W1 x + b1) + b2) + b3
// the input gradient is g = W1^T (m1 * (W2^T (m2 * w3))) with m1, m2 the ReLU masks; the penalty
// 20*mean((|g|-1)^2) is differentiated through that chain with the masks held constant (ReLU has zero
// second derivative), which is exactly what autograd's create_graph=True double backward computes.
// The "zero diff" positive sample D(0) is carried as one extra row (row M) of the minibatch.
#include "common.cuh"
#include "addk.h"
#include <stdlib.h>
#include "switches.h"
#include "h3_scale.cuh"
#include "adam.cuh"
#include <cuda_fp16.h>

struct addk_update_ctx {
#define ADDK_PTR(n) void* n;
#define ADDK_INT(n) int64_t n;
#define ADDK_F64(n) double n;
#include "ctx_fields.h"
};

int addk_gemm_tc(cudaStream_t st, const addk_gemm_args& a, int precision);
bool addk_gemm_h3_usable(const addk_gemm_args& a);
namespace addk { int sgemm_launch(cudaStream_t st, const addk_gemm_args& a); }

namespace addk {

enum { ST_SURR = 0, ST_CLIP, ST_RATIO, ST_BOUND, ST_CRITIC, ST_BCE_NEG, ST_BCE_POS, ST_PEN, ST_NEG_LOGIT,
       ST_POS_LOGIT, ST_NEG_ACC, ST_POS_ACC, ST_WL_SQ, ST_W_SQ, ST_COUNT = 32 };

// precision "f16x3": an elementwise kernel that produces a dense-layer operand also leaves max|x| of what it wrote in
// the operand's slot (word [1]), so the conversion that follows is the split pass alone.  EVERY thread of the block
// calls this once with its own maximum (no early returns before it): block reduction, then one thread looks at the slot
// and issues the atomic only if it would raise it.  (A first version did this per warp: 262k warps polling one L2
// address made the optimizer step 3 % slower than the separate max passes it replaced.)
__device__ __forceinline__ void amax_commit(float v, uint32_t* slot) {
  if (!slot) return;                                   // uniform across the grid
  __shared__ uint32_t s_m[32];
  uint32_t m = __reduce_max_sync(0xffffffffu, __float_as_uint(v));
  if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? s_m[threadIdx.x] : 0u;
    m = __reduce_max_sync(0xffffffffu, m);
    if (threadIdx.x == 0 && m > *reinterpret_cast<volatile uint32_t*>(slot + 1)) atomicMax(slot + 1, m);
  }
  __syncthreads();                                     // s_m is reused by the next call
}

// ---- ExperienceBuffer.sample + normalisation: one warp per minibatch row ---------------------------------
__global__ void gather_minibatch_kernel(const long long* __restrict__ idx, int M, int obs_dim, int obs_ld, int act_dim, int act_ld,
                                        int disc_dim, int disc_ld, const float* __restrict__ buf_obs,
                                        const float* __restrict__ buf_action, const float* __restrict__ buf_logp,
                                        const float* __restrict__ buf_adv, const float* __restrict__ buf_tar,
                                        const float* __restrict__ buf_mask, const float* __restrict__ buf_dobs,
                                        const float* __restrict__ buf_demo, const float* __restrict__ obs_mean,
                                        const float* __restrict__ obs_std, const float* __restrict__ a_mean,
                                        const float* __restrict__ a_std, const float* __restrict__ mean_abs,
                                        float* __restrict__ xn, float* __restrict__ an, float* __restrict__ old_logp,
                                        float* __restrict__ adv, float* __restrict__ tar, float* __restrict__ mask,
                                        float* __restrict__ dn, int* __restrict__ cnt, uint16_t* __restrict__ xn16,
                                        uint16_t* __restrict__ dn16, uint32_t* xn_slot, uint32_t* dn_slot) {
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  const int i = blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
  const int lane = threadIdx.x & 31;
  float vx = 0.f, vd = 0.f;                              // max|xn|, max|dn| of what this thread writes
  if (i < M) {
    const size_t s = (size_t)idx[i];
    // same arithmetic per element as before ((x - mean) / std with a separate subtraction), 128- / 64-bit accesses where
    // the row geometry allows: obs rows are 16-byte aligned when obs_dim % 4 == 0, disc rows 8-byte when disc_dim % 2 == 0
    if ((obs_dim & 3) == 0 && (obs_ld & 3) == 0) {
      const float* src = buf_obs + s * obs_dim;
      float* dst = xn + (size_t)i * obs_ld;
      for (int c = 4 * lane; c < obs_ld; c += 128) {        // columns obs_dim .. obs_ld-1 are alignment padding (zeros)
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < obs_dim) {
          const float4 v = ldg4(src + c), m = ldg4(obs_mean + c), d = ldg4(obs_std + c);
          o = make_float4(sub_rn(v.x, m.x) / d.x, sub_rn(v.y, m.y) / d.y, sub_rn(v.z, m.z) / d.z, sub_rn(v.w, m.w) / d.w);
        }
        stg4(dst + c, o);
        vx = fmaxf(fmaxf(vx, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
        if (xn16) {
          uint16_t* d16 = xn16 + (size_t)i * obs_ld + c;
          d16[0] = to_bf16(o.x); d16[1] = to_bf16(o.y); d16[2] = to_bf16(o.z); d16[3] = to_bf16(o.w);
        }
      }
    } else {
      for (int c = lane; c < obs_ld; c += 32) {
        const float v = c < obs_dim ? sub_rn(buf_obs[s * obs_dim + c], obs_mean[c]) / obs_std[c] : 0.f;
        xn[(size_t)i * obs_ld + c] = v;
        vx = fmaxf(vx, fabsf(v));
        if (xn16) xn16[(size_t)i * obs_ld + c] = to_bf16(v);
      }
    }
    for (int c = lane; c < act_ld; c += 32)
      an[(size_t)i * act_ld + c] = c < act_dim ? sub_rn(buf_action[s * act_dim + c], a_mean[c]) / a_std[c] : 0.f;
    if ((disc_dim & 1) == 0 && (disc_ld & 1) == 0) {
      const float* pd = buf_demo + s * disc_dim;
      const float* po = buf_dobs + s * disc_dim;
      float* dst = dn + (size_t)i * disc_ld;
      for (int c = 2 * lane; c < disc_ld; c += 64) {
        float2 o = make_float2(0.f, 0.f);
        if (c < disc_dim) {
          const float2 a = __ldg(reinterpret_cast<const float2*>(pd + c)), b = __ldg(reinterpret_cast<const float2*>(po + c));
          const float2 m = __ldg(reinterpret_cast<const float2*>(mean_abs + c));
          o = make_float2(sub_rn(a.x, b.x) / fmaxf(m.x, 1e-4f), sub_rn(a.y, b.y) / fmaxf(m.y, 1e-4f));
        }
        *reinterpret_cast<float2*>(dst + c) = o;
        vd = fmaxf(vd, fmaxf(fabsf(o.x), fabsf(o.y)));
        if (dn16) { dn16[(size_t)i * disc_ld + c] = to_bf16(o.x); dn16[(size_t)i * disc_ld + c + 1] = to_bf16(o.y); }
      }
    } else {
      for (int c = lane; c < disc_ld; c += 32) {
        const float v = c < disc_dim ? sub_rn(buf_demo[s * disc_dim + c], buf_dobs[s * disc_dim + c]) / fmaxf(mean_abs[c], 1e-4f) : 0.f;
        dn[(size_t)i * disc_ld + c] = v;
        vd = fmaxf(vd, fabsf(v));
        if (dn16) dn16[(size_t)i * disc_ld + c] = to_bf16(v);
      }
    }
    if (lane == 0) {
      old_logp[i] = buf_logp[s]; adv[i] = buf_adv[s]; tar[i] = buf_tar[s];
      const float mk = buf_mask[s];
      mask[i] = mk;
      if (mk == 1.0f) atomicAdd(&s_cnt, 1);
    }
  }
  amax_commit(vx, xn_slot);
  amax_commit(vd, dn_slot);
  __syncthreads();
  if (threadIdx.x == 0 && s_cnt) atomicAdd(cnt, s_cnt);       // one global atomic per block instead of one per row
}

__device__ __forceinline__ float gaussian_logp(float sq_sum, float logstd_sum, int dim) {
  // -0.5*sum((x-mean)/std)^2  +  (-0.5*dim*log(2*pi) - sum(logstd))   distribution_gaussian_diag.py:90-94
  const float c = (float)(-0.5 * (double)dim * 1.8378770664093453);
  return add_rn(mul_rn(-0.5f, sq_sum), sub_rn(c, logstd_sum));
}

// ---- PPO surrogate + action-bound loss and d(loss)/d(mean): one warp per row ------------------------------
__global__ void actor_loss_kernel(const float* __restrict__ mean, const float* __restrict__ an,
                                  const float* __restrict__ logstd, const float* __restrict__ old_logp,
                                  const float* __restrict__ adv, const float* __restrict__ mask, int M, int act_dim,
                                  int act_ld, float clip, float bound_w, const int* __restrict__ cnt,
                                  float* __restrict__ dmean, double* __restrict__ stats, uint16_t* __restrict__ dmean16,
                                  uint32_t* dmean_slot) {
  __shared__ double s_acc[8][4];
  const int warp = threadIdx.x / 32, lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x / 32) + warp;
  double a_surr = 0.0, a_clip = 0.0, a_ratio = 0.0, a_bound = 0.0;
  float gabs = 0.f;
  if (i < M) {
    const bool on = mask[i] == 1.0f;
    float m = 0.f, d = 0.f, sd = 1.f, ls = 0.f;
    if (lane < act_dim) {
      m = mean[(size_t)i * act_ld + lane];
      ls = logstd[lane];
      sd = expf(ls);
      d = sub_rn(an[(size_t)i * act_ld + lane], m);
    }
    float z = d / sd;
    float sq = warp_sum(lane < act_dim ? mul_rn(z, z) : 0.f);
    float lss = warp_sum(ls);
    float logp = gaussian_logp(sq, lss, act_dim);
    float ratio = expf(sub_rn(logp, old_logp[i]));
    float a = adv[i];
    float l0 = mul_rn(a, ratio);
    float rc = fminf(fmaxf(ratio, 1.0f - clip), 1.0f + clip);
    float l1 = mul_rn(a, rc);
    float surr = fminf(l0, l1);
    // d surr / d ratio: through l0 when l0 <= l1, through l1 only inside the clip range (ties split 1/2+1/2)
    float dr;
    const bool inside = ratio >= 1.0f - clip && ratio <= 1.0f + clip;
    if (l0 < l1) dr = a;
    else if (l0 == l1) dr = 0.5f * a + (inside ? 0.5f * a : 0.f);
    else dr = inside ? a : 0.f;
    const float n = (float)max(*cnt, 1);
    float dlogp = -(dr * ratio) / n;   // actor_loss = -mean(surr)
    float vmin = fminf(add_rn(m, 1.0f), 0.f), vmax = fmaxf(sub_rn(m, 1.0f), 0.f);
    float viol = lane < act_dim ? add_rn(mul_rn(vmin, vmin), mul_rn(vmax, vmax)) : 0.f;
    float viol_sum = warp_sum(viol);
    if (lane < act_ld) {
      float g = 0.f;
      if (on && lane < act_dim) g = dlogp * (d / (sd * sd)) + bound_w * 2.0f * (vmin + vmax) / n;
      dmean[(size_t)i * act_ld + lane] = g;
      gabs = fabsf(g);
      if (dmean16) dmean16[(size_t)i * act_ld + lane] = to_bf16(g);
    }
    if (on) {
      a_surr = surr; a_clip = fabsf(sub_rn(ratio, 1.0f)) > clip ? 1.0 : 0.0; a_ratio = ratio; a_bound = viol_sum;
    }
  }
  amax_commit(gabs, dmean_slot);
  if (lane == 0) { s_acc[warp][0] = a_surr; s_acc[warp][1] = a_clip; s_acc[warp][2] = a_ratio; s_acc[warp][3] = a_bound; }
  __syncthreads();
  if (threadIdx.x < 4) {          // one atomic per statistic per block instead of one per row
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x / 32); ++w) t += s_acc[w][threadIdx.x];
    const int slot = threadIdx.x == 0 ? ST_SURR : threadIdx.x == 1 ? ST_CLIP : threadIdx.x == 2 ? ST_RATIO : ST_BOUND;
    atomicAdd(stats + slot, t);
  }
}

__global__ void critic_loss_kernel(const float* __restrict__ pred, const float* __restrict__ tar, int M, float w,
                                   float* __restrict__ dpred, double* __restrict__ stats) {
  __shared__ double sm[32];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double s = 0.0;
  if (i < M) {
    float diff = sub_rn(tar[i], pred[i]);
    s = (double)diff * diff;
    dpred[i] = w * (-2.0f * diff) / (float)M;
  }
  s = block_sum(s, sm);
  if (threadIdx.x == 0) atomicAdd(stats + ST_CRITIC, s);
}

// BCEWithLogits with smoothed labels 0.1 (rows < M, mean over M) and 0.9 (row M): amp_agent.py:177-185
__global__ void disc_loss_kernel(const float* __restrict__ logit, int M, float w, float* __restrict__ dlogit,
                                 double* __restrict__ stats) {
  __shared__ double sm[32];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double bce_n = 0.0, lg = 0.0, acc = 0.0;
  if (i <= M) {
    float x = logit[i];
    float t = i < M ? 0.1f : 0.9f;
    float bce = fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
    float sg = 1.0f / (1.0f + expf(-x));
    float wt = i < M ? 1.0f / (float)M : 1.0f;
    dlogit[i] = w * 0.5f * (sg - t) * wt;
    if (i < M) { bce_n = bce; lg = x; acc = x < 0.f ? 1.0 : 0.0; }
    else {
      atomicAdd(stats + ST_BCE_POS, (double)bce);
      atomicAdd(stats + ST_POS_LOGIT, (double)x);
      atomicAdd(stats + ST_POS_ACC, x > 0.f ? 1.0 : 0.0);
    }
  }
  bce_n = block_sum(bce_n, sm); lg = block_sum(lg, sm); acc = block_sum(acc, sm);
  if (threadIdx.x == 0) {
    atomicAdd(stats + ST_BCE_NEG, bce_n); atomicAdd(stats + ST_NEG_LOGIT, lg); atomicAdd(stats + ST_NEG_ACC, acc);
  }
}

// u2 = relu'(h2) * w3 (gradient of the logit w.r.t. the last hidden layer) and dh2 = dlogit * u2
__global__ void disc_head_backward_kernel(const float* __restrict__ h2, const float* __restrict__ wl,
                                          const float* __restrict__ dlogit, int R, int H, float* __restrict__ u2,
                                          float* __restrict__ dh2, uint16_t* __restrict__ u2_16, uint16_t* __restrict__ dh2_16,
                                          uint32_t* u2_slot, uint32_t* dh2_slot) {
  float um = 0.f, dm = 0.f;
  const size_t tot = (size_t)R * H;
  if ((H & 3) == 0 && !u2_16 && !dh2_16) {          // 128-bit path (f16x3 / fp32 modes)
    const size_t n4 = tot >> 2;
    const int h4 = H >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
      const int r = (int)(i / h4), k = (int)(i - (size_t)r * h4) * 4;
      const float4 h = ldg4(h2 + 4 * i), w = ldg4(wl + k);
      const float dl = dlogit[r];
      const float4 u = make_float4(h.x > 0.f ? w.x : 0.f, h.y > 0.f ? w.y : 0.f, h.z > 0.f ? w.z : 0.f, h.w > 0.f ? w.w : 0.f);
      const float4 d = make_float4(dl * u.x, dl * u.y, dl * u.z, dl * u.w);
      stg4(u2 + 4 * i, u);
      stg4(dh2 + 4 * i, d);
      um = fmaxf(fmaxf(um, fmaxf(fabsf(u.x), fabsf(u.y))), fmaxf(fabsf(u.z), fabsf(u.w)));
      dm = fmaxf(fmaxf(dm, fmaxf(fabsf(d.x), fabsf(d.y))), fmaxf(fabsf(d.z), fabsf(d.w)));
    }
    amax_commit(um, u2_slot);
    amax_commit(dm, dh2_slot);
    return;
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / H), k = (int)(i - (size_t)r * H);
    const float u = h2[i] > 0.f ? wl[k] : 0.f;
    u2[i] = u;
    const float dh = dlogit[r] * u;
    dh2[i] = dh;
    if (u2_16) u2_16[i] = to_bf16(u);
    if (dh2_16) dh2_16[i] = to_bf16(dh);
    um = fmaxf(um, fabsf(u)); dm = fmaxf(dm, fabsf(dh));
  }
  amax_commit(um, u2_slot);
  amax_commit(dm, dh2_slot);
}

// gradient penalty on the input gradient: one warp per row (add_agent.py:167-178)
__global__ void grad_penalty_kernel(const float* __restrict__ gx, int M, int R, int dim, int ld, float coef,
                                    float* __restrict__ dg, double* __restrict__ stats, uint16_t* __restrict__ dg16,
                                    uint32_t* dg_slot) {
  __shared__ double s_pen[8];
  int i = blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
  int lane = threadIdx.x & 31;
  float vm = 0.f;
  double pen = 0.0;
  if (i < R) {                                        // whole warps
    float s = 0.f;
    for (int c = lane; c < dim; c += 32) { float v = gx[(size_t)i * ld + c]; s += v * v; }
    s = warp_sum(s);
    float gn = sqrtf(s + 1e-8f);
    float e = gn - 1.0f;
    float sc = i < M ? coef * 2.0f * e / (gn * (float)M) : 0.f;
    for (int c = lane; c < ld; c += 32) {
      const float v = c < dim ? sc * gx[(size_t)i * ld + c] : 0.f;
      dg[(size_t)i * ld + c] = v;
      vm = fmaxf(vm, fabsf(v));
      if (dg16) dg16[(size_t)i * ld + c] = to_bf16(v);
    }
    if (i < M) pen = (double)e * e;
  }
  amax_commit(vm, dg_slot);
  // one atomic per block (16384 double atomics on ONE address cost 30 us of a 33 us kernel); the 8 rows of a block are
  // added in a fixed order, the blocks in arrival order as before
  if (lane == 0) s_pen[threadIdx.x >> 5] = pen;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_pen[w];
    if (t != 0.0) atomicAdd(stats + ST_PEN, t);
  }
}

// sum of squares of the discriminator's three weight tensors in one launch: blockIdx.y selects the tensor;
// y = 0 (logit weights) feeds both statistics (disc_logit_reg and disc_weight_decay terms, add_agent.py:157-202)
__global__ void disc_weight_sumsq_kernel(const float* __restrict__ wl, long long nl, const float* __restrict__ w0, long long n0,
                                         const float* __restrict__ w1, long long n1, double* __restrict__ stats) {
  __shared__ double sm[32];
  const float* x = blockIdx.y == 0 ? wl : (blockIdx.y == 1 ? w0 : w1);
  const long long n = blockIdx.y == 0 ? nl : (blockIdx.y == 1 ? n0 : n1);
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = x[i]; s += v * v;
  }
  s = block_sum(s, sm);
  if (threadIdx.x == 0 && s != 0.0) {
    atomicAdd(stats + ST_W_SQ, s);
    if (blockIdx.y == 0) atomicAdd(stats + ST_WL_SQ, s);
  }
}

__global__ void sumsq_kernel(const float* __restrict__ x, long long n, double* __restrict__ out) {
  __shared__ double sm[32];
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = x[i]; s += v * v;
  }
  s = block_sum(s, sm);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

// Column sums over the minibatch rows: out[n] = sum_r (rw ? rw[r] : 1) * dY[r, n]  -- bias gradients (rw = NULL) and
// the weight gradient of a 1-output head (rw = d loss / d output).  HBM-bound (reads rows x n floats once).
// grid = (ceil(n/128), COLSUM_CHUNKS): every block reduces a 128-column x rows/64 patch into `work`; the last block of
// a column group to finish (ticket counter, self-resetting) adds the 64 partials in a fixed order, writes the total
// to slab 0 and zeroes slabs 1..nsplit-1 so the common slab reduction stays valid.  Deterministic, no float atomics.
constexpr int COLSUM_CHUNKS = 128;      // rows of the work buffer = the largest gridDim.y (the launch picks ~1000 blocks)
constexpr int COLSUM_MAX_N = 1024;
// IN16: dY is read from its bf16 copy (a tensor that has no fp32 copy, precision "bf16")
__device__ __forceinline__ float4 ld_row4(const float* p) { return ldg4(p); }
__device__ __forceinline__ float4 ld_row4(const uint16_t* p) {
  const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
  return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xFFFF0000u), __uint_as_float(r.y << 16),
                     __uint_as_float(r.y & 0xFFFF0000u));
}
__device__ __forceinline__ float ld_row1(const float* p) { return *p; }
__device__ __forceinline__ float ld_row1(const uint16_t* p) { return __uint_as_float((uint32_t)*p << 16); }
// precision "f16x3": a tensor that exists only as its fp16 planes: x = (hi + lo) / s   (22+ bits, see gemm_tc.cu)
struct PlanesPtr { const __half* hi; long long plane; float inv; };
__device__ __forceinline__ PlanesPtr operator+(const PlanesPtr& p, size_t off) { return PlanesPtr{p.hi + off, p.plane, p.inv}; }
__device__ __forceinline__ float4 ld_row4(const PlanesPtr& p) {
  const uint2 h = __ldg(reinterpret_cast<const uint2*>(p.hi)), l = __ldg(reinterpret_cast<const uint2*>(p.hi + p.plane));
  const float2 h01 = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), h23 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
  const float2 l01 = __half22float2(*reinterpret_cast<const __half2*>(&l.x)), l23 = __half22float2(*reinterpret_cast<const __half2*>(&l.y));
  return make_float4((h01.x + l01.x) * p.inv, (h01.y + l01.y) * p.inv, (h23.x + l23.x) * p.inv, (h23.y + l23.y) * p.inv);
}
__device__ __forceinline__ float ld_row1(const PlanesPtr& p) { return (__half2float(*p.hi) + __half2float(p.hi[p.plane])) * p.inv; }
template <typename T> struct ColIn { typedef const T* type; };
template <> struct ColIn<PlanesPtr> { typedef PlanesPtr type; };
struct PlanesArg { const __half* hi; long long plane; const uint32_t* slot; };

template <typename TIN>
__global__ void __launch_bounds__(256) colsum_slabs_kernel(const TIN* __restrict__ dY_raw, PlanesArg pl, int ld, int rows, int n,
                                                           float* __restrict__ out, long long slab_stride, int nsplit,
                                                           const float* __restrict__ rw, float* __restrict__ work,
                                                           unsigned int* __restrict__ tickets) {
  __shared__ float4 sm[8][32];
  __shared__ unsigned int s_last;
  typename ColIn<TIN>::type dY;
  if constexpr (sizeof(TIN) == sizeof(PlanesPtr)) {
    float sc, inv;
    addk_tc::h3_slot_scale(pl.slot, sc, inv);
    dY = PlanesPtr{pl.hi, pl.plane, inv};
  } else {
    dY = dY_raw;
  }
  const int cq = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cq) * 4;
  const int chunks = (int)gridDim.y;
  const int per = (rows + chunks - 1) / chunks;
  const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < n) {
    if (col + 3 < n && (ld & 3) == 0) {
      int r = r0 + rl;
      // eight, then four independent 128-bit loads in flight per thread (64 chunks x 4 loads: 2.5 TB/s on a 64 MB tensor)
      for (; r + 56 < r1; r += 64) {
        float4 v[8];
        float w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { v[u] = ld_row4(dY + (size_t)(r + 8 * u) * ld + col); w[u] = rw ? rw[r + 8 * u] : 1.f; }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          acc.x = fmaf(w[u], v[u].x, acc.x); acc.y = fmaf(w[u], v[u].y, acc.y); acc.z = fmaf(w[u], v[u].z, acc.z); acc.w = fmaf(w[u], v[u].w, acc.w);
        }
      }
      for (; r + 24 < r1; r += 32) {
        const float4 v0 = ld_row4(dY + (size_t)r * ld + col), v1 = ld_row4(dY + (size_t)(r + 8) * ld + col);
        const float4 v2 = ld_row4(dY + (size_t)(r + 16) * ld + col), v3 = ld_row4(dY + (size_t)(r + 24) * ld + col);
        const float w0 = rw ? rw[r] : 1.f, w1 = rw ? rw[r + 8] : 1.f, w2 = rw ? rw[r + 16] : 1.f, w3 = rw ? rw[r + 24] : 1.f;
        acc.x = fmaf(w0, v0.x, acc.x); acc.y = fmaf(w0, v0.y, acc.y); acc.z = fmaf(w0, v0.z, acc.z); acc.w = fmaf(w0, v0.w, acc.w);
        acc.x = fmaf(w1, v1.x, acc.x); acc.y = fmaf(w1, v1.y, acc.y); acc.z = fmaf(w1, v1.z, acc.z); acc.w = fmaf(w1, v1.w, acc.w);
        acc.x = fmaf(w2, v2.x, acc.x); acc.y = fmaf(w2, v2.y, acc.y); acc.z = fmaf(w2, v2.z, acc.z); acc.w = fmaf(w2, v2.w, acc.w);
        acc.x = fmaf(w3, v3.x, acc.x); acc.y = fmaf(w3, v3.y, acc.y); acc.z = fmaf(w3, v3.z, acc.z); acc.w = fmaf(w3, v3.w, acc.w);
      }
      for (; r < r1; r += 8) {
        const float4 v = ld_row4(dY + (size_t)r * ld + col);
        const float w = rw ? rw[r] : 1.f;
        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
      }
    } else {
      for (int r = r0 + rl; r < r1; r += 8) {
        const typename ColIn<TIN>::type q = dY + (size_t)r * ld + col;
        const float w = rw ? rw[r] : 1.f;
        acc.x = fmaf(w, ld_row1(q), acc.x);
        if (col + 1 < n) acc.y = fmaf(w, ld_row1(q + 1), acc.y);
        if (col + 2 < n) acc.z = fmaf(w, ld_row1(q + 2), acc.z);
        if (col + 3 < n) acc.w = fmaf(w, ld_row1(q + 3), acc.w);
      }
    }
  }
  sm[rl][cq] = acc;
  __syncthreads();
  if (rl == 0) {
    float4 t = sm[0][cq];
#pragma unroll
    for (int i = 1; i < 8; ++i) { t.x += sm[i][cq].x; t.y += sm[i][cq].y; t.z += sm[i][cq].z; t.w += sm[i][cq].w; }
    *reinterpret_cast<float4*>(work + ((size_t)blockIdx.y * COLSUM_MAX_N + blockIdx.x * 128 + cq * 4)) = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(tickets + blockIdx.x, 1u) == (unsigned)chunks - 1u) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int ch = rl; ch < chunks; ch += 8) {
    const float4 v = __ldcg(reinterpret_cast<const float4*>(work + ((size_t)ch * COLSUM_MAX_N + blockIdx.x * 128 + cq * 4)));
    t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
  }
  sm[rl][cq] = t;
  __syncthreads();
  if (rl == 0 && col < n) {
    t = sm[0][cq];
#pragma unroll
    for (int i = 1; i < 8; ++i) { t.x += sm[i][cq].x; t.y += sm[i][cq].y; t.z += sm[i][cq].z; t.w += sm[i][cq].w; }
    const float tv[4] = {t.x, t.y, t.z, t.w};
    for (int e = 0; e < 4 && col + e < n; ++e) {
      out[col + e] = tv[e];
      for (int z = 1; z < nsplit; ++z) out[(size_t)z * slab_stride + col + e] = 0.f;
    }
  }
  if (threadIdx.x == 0) tickets[blockIdx.x] = 0u;
}

// 1-output head forward: out[r] = dot(X[r, :], w) + b   (critic value, discriminator logit); warp per row
__global__ void rowdot_kernel(const float* __restrict__ X, int ld, long long rows, int K, const float* __restrict__ w,
                              const float* __restrict__ b, float* __restrict__ out) {
  const long long r = (long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* x = X + r * ld;
  float acc = 0.f;
  if ((K & 3) == 0 && (ld & 3) == 0) {
    for (int k = lane * 4; k < K; k += 128) {
      const float4 v = ldg4(x + k), u = ldg4(w + k);
      acc = fmaf(v.x, u.x, acc); acc = fmaf(v.y, u.y, acc); acc = fmaf(v.z, u.z, acc); acc = fmaf(v.w, u.w, acc);
    }
  } else {
    for (int k = lane; k < K; k += 32) acc = fmaf(x[k], w[k], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[r] = acc + (b ? b[0] : 0.f);
}

// 1-output head input gradient through the ReLU: g[r, k] = (h[r, k] > 0) ? d[r] * w[k] : 0
__global__ void outer_mask_kernel(const float* __restrict__ d, const float* __restrict__ w, const float* __restrict__ h,
                                  long long rows, int K, float* __restrict__ g, uint16_t* __restrict__ g16, uint32_t* g_slot) {
  const long long n4 = rows * K / 4;
  float vm = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = (i * 4) / K;
    const int k = (int)((i * 4) % K);
    const float dv = d[r];
    const float4 u = ldg4(w + k), m = ldg4(h + 4 * i);
    const float4 o = make_float4(m.x > 0.f ? dv * u.x : 0.f, m.y > 0.f ? dv * u.y : 0.f, m.z > 0.f ? dv * u.z : 0.f,
                                 m.w > 0.f ? dv * u.w : 0.f);
    stg4(g + 4 * i, o);
    vm = fmaxf(fmaxf(vm, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
    if (g16) { g16[4 * i] = to_bf16(o.x); g16[4 * i + 1] = to_bf16(o.y); g16[4 * i + 2] = to_bf16(o.z); g16[4 * i + 3] = to_bf16(o.w); }
  }
  amax_commit(vm, g_slot);
}

// out[n] = sum over the `parts` per-block partial rows the f16x3 split pass left behind (h3_split_kernel), in a fixed
// order; slabs 1..nsplit-1 are zeroed like colsum_slabs_kernel does.  grid = (ceil(n / 32), CP_CHUNKS): every block adds
// its chunk of the partial rows for 8 float4 columns x 32 row lanes (four independent loads in flight per thread) and
// leaves one row in `work`; the last block of a column group to finish (ticket counter, self-resetting) adds the
// CP_CHUNKS rows in a fixed order.  (One block per column group walked up to 1184 rows with 16 - 32 blocks in the whole
// grid: 10 us per launch, eight launches per optimizer step.)
constexpr int CP_CHUNKS = 8;
__global__ void __launch_bounds__(256) colsum_parts_kernel(const float* __restrict__ part, int parts, int n, float* __restrict__ out,
                                                           long long slab_stride, int nsplit, float* __restrict__ work,
                                                           unsigned int* __restrict__ tickets) {
  __shared__ float4 sm[32][8];
  __shared__ unsigned int s_last;
  const int cq = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int col = (blockIdx.x * 8 + cq) * 4;
  const int per = (parts + CP_CHUNKS - 1) / CP_CHUNKS;
  const int r_begin = blockIdx.y * per, r_end = min(parts, r_begin + per);
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < n) {
    float4 u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) u[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = r_begin + rl; r < r_end; r += 128) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rr = r + 32 * j;
        if (rr < r_end) {
          const float4 v = ldg4(part + (size_t)rr * n + col);
          u[j].x += v.x; u[j].y += v.y; u[j].z += v.z; u[j].w += v.w;
        }
      }
    }
    t.x = (u[0].x + u[1].x) + (u[2].x + u[3].x); t.y = (u[0].y + u[1].y) + (u[2].y + u[3].y);
    t.z = (u[0].z + u[1].z) + (u[2].z + u[3].z); t.w = (u[0].w + u[1].w) + (u[2].w + u[3].w);
  }
  sm[rl][cq] = t;
  __syncthreads();
  if (rl == 0) {
    t = sm[0][cq];
#pragma unroll
    for (int i = 1; i < 32; ++i) { t.x += sm[i][cq].x; t.y += sm[i][cq].y; t.z += sm[i][cq].z; t.w += sm[i][cq].w; }
    *reinterpret_cast<float4*>(work + ((size_t)blockIdx.y * COLSUM_MAX_N + blockIdx.x * 32 + cq * 4)) = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(tickets + blockIdx.x, 1u) == CP_CHUNKS - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (rl == 0 && col < n) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ch = 0; ch < CP_CHUNKS; ++ch) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(work + ((size_t)ch * COLSUM_MAX_N + blockIdx.x * 32 + cq * 4)));
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    const float tv[4] = {a.x, a.y, a.z, a.w};
    for (int e = 0; e < 4 && col + e < n; ++e) {
      out[col + e] = tv[e];
      for (int z = 1; z < nsplit; ++z) out[(size_t)z * slab_stride + col + e] = 0.f;
    }
  }
  if (threadIdx.x == 0) tickets[blockIdx.x] = 0u;
}

// torch.nn.utils.clip_grad_norm_ (mp_optimizer.py:19-20,46-47): total = ||g||_2 over all parameters,
// g *= min(1, max_norm / (total + 1e-6)).  `pre_scale` = 1/world when g still holds the cross-rank SUM.
__global__ void clip_coef_kernel(const double* __restrict__ sumsq, float pre_scale, float max_norm, float* __restrict__ coef) {
  const float total = pre_scale * (float)sqrt(*sumsq);
  *coef = fminf(max_norm / (total + 1e-6f), 1.0f);
}
__global__ void scale_by_coef_kernel(float* __restrict__ g, long long n, const float* __restrict__ coef) {
  const float c = *coef;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<float4*>(g)[i];
    v.x *= c; v.y *= c; v.z *= c; v.w *= c;
    reinterpret_cast<float4*>(g)[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) g[4 * n4 + threadIdx.x] *= c;
}

struct InfoArgs { const double* st; const int* cnt; int M; float bound_w, critic_w, disc_w, logit_reg, gp, wd; float* info; };
__device__ __forceinline__ void finalize_info(const InfoArgs& a) {
  const double* st = a.st;
  const int M = a.M;
  const float bound_w = a.bound_w, critic_w = a.critic_w, disc_w = a.disc_w, logit_reg = a.logit_reg, gp = a.gp, wd = a.wd;
  float* info = a.info;
  double n = (double)max(*a.cnt, 1);
  double surr = st[ST_SURR] / n, bound = st[ST_BOUND] / n;
  double actor = -surr + (bound_w != 0.f ? bound_w * bound : 0.0);
  double critic = st[ST_CRITIC] / M;
  double pen = st[ST_PEN] / M;
  double disc = 0.5 * (st[ST_BCE_POS] + st[ST_BCE_NEG] / M) + logit_reg * st[ST_WL_SQ] + gp * pen + wd * st[ST_W_SQ];
  info[0] = (float)(actor + critic_w * critic + disc_w * disc);
  info[1] = (float)critic;
  info[2] = (float)actor;
  info[3] = (float)(st[ST_CLIP] / n);
  info[4] = (float)(st[ST_RATIO] / n);
  info[5] = (float)bound;
  info[6] = (float)disc;
  info[7] = (float)pen;
  info[8] = (float)st[ST_WL_SQ];
  info[9] = (float)st[ST_POS_ACC];
  info[10] = (float)(st[ST_NEG_ACC] / M);
  info[11] = (float)st[ST_POS_LOGIT];
  info[12] = (float)(st[ST_NEG_LOGIT] / M);
  info[13] = (float)n;
  info[14] = 0.f; info[15] = 0.f;
}
struct Seg { long long begin, end; int nslabs; float l2; };
struct SegTable { Seg s[24]; int n; long long P; };

// grads[i] = sum over the split-K slabs of segment(i) + l2 * param[i]
__global__ void reduce_slabs_kernel(const __grid_constant__ SegTable t, const float* __restrict__ slabs,
                                    const float* __restrict__ params, float* __restrict__ grads, const InfoArgs ia) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && ia.info) finalize_info(ia);      // the diagnostics row of the step
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= t.P) return;
  int ns = 0; float l2 = 0.f;
  for (int k = 0; k < t.n; ++k) if (i >= t.s[k].begin && i < t.s[k].end) { ns = t.s[k].nslabs; l2 = t.s[k].l2; break; }
  float g = 0.f;
  for (int s = 0; s < ns; ++s) g += slabs[(size_t)s * t.P + i];
  if (l2 != 0.f) g += l2 * params[i];
  grads[i] = g;
}


// The tail of an optimizer step as ONE launch (single GPU, no gradient clipping): slab reduction + AdamW + the
// diagnostics row.  Same arithmetic, element by element, as reduce_slabs_kernel (slab sums + diagnostics row) followed by
// adamw_vec4_kernel (adam1); the summed gradient is still written (diagnostics / tests read it).  As three launches behind
// the join of the chains the tail ran alone on the GPU for ~80 us of a 1.7 ms step; fused it reads the slabs once and
// never re-reads the gradient vector.
__global__ void __launch_bounds__(256) reduce_slabs_adamw_kernel(const __grid_constant__ SegTable t, const float* __restrict__ slabs,
                                                                  float* __restrict__ params, float* __restrict__ grads,
                                                                  float* __restrict__ m, float* __restrict__ v, const AdamK k,
                                                                  const InfoArgs ia) {
  if (blockIdx.x == 0 && threadIdx.x == 0) finalize_info(ia);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= t.P) return;
  int ns = 0; float l2 = 0.f;
  for (int q = 0; q < t.n; ++q) if (i >= t.s[q].begin && i < t.s[q].end) { ns = t.s[q].nslabs; l2 = t.s[q].l2; break; }
  float pi = params[i], mi = m[i], vi = v[i];
  const float* sp = slabs + i;
  float x[8];
  float g = 0.f;
  int s = 0;
  for (; s + 8 <= ns; s += 8) {           // eight slab loads in flight, added in slab order
#pragma unroll
    for (int u = 0; u < 8; ++u) x[u] = __ldcs(sp + (size_t)(s + u) * t.P);
#pragma unroll
    for (int u = 0; u < 8; ++u) g += x[u];
  }
  for (; s < ns; ++s) g += __ldcs(sp + (size_t)s * t.P);
  if (l2 != 0.f) g += l2 * pi;
  grads[i] = g;
  adam1(k, pi, g, mi, vi);
  params[i] = pi; m[i] = mi; v[i] = vi;
}

// DistributionGaussianDiag.sample/log_prob + rand_action_mask select + Normalizer.unnormalize: warp per env
__global__ void sample_action_kernel(const float* __restrict__ mean, int act_ld, const float* __restrict__ logstd,
                                     const float* __restrict__ noise, const float* __restrict__ exp_mask,
                                     const float* __restrict__ a_mean, const float* __restrict__ a_std, int n,
                                     int act_dim, float* __restrict__ action, float* __restrict__ a_logp,
                                     float* __restrict__ action_rec, float* __restrict__ logp_rec,
                                     float* __restrict__ mask_rec) {
  int i = blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
  int lane = threadIdx.x & 31;
  if (i >= n) return;
  const float mk = exp_mask ? exp_mask[i] : 1.0f;
  float m = 0.f, ls = 0.f, z = 0.f;
  if (lane < act_dim) {
    m = mean[(size_t)i * act_ld + lane];
    ls = logstd[lane];
    float sd = expf(ls);
    float x = add_rn(m, mul_rn(sd, noise[(size_t)i * act_dim + lane]));
    float na = (mk == 1.0f) ? x : m;
    z = sub_rn(na, m) / sd;
    float a = add_rn(mul_rn(na, a_std[lane]), a_mean[lane]);
    action[(size_t)i * act_dim + lane] = a;
    if (action_rec) action_rec[(size_t)i * act_dim + lane] = a;
  }
  float sq = warp_sum(lane < act_dim ? mul_rn(z, z) : 0.f);
  float lss = warp_sum(ls);
  if (lane == 0) {
    float lp = gaussian_logp(sq, lss, act_dim);
    a_logp[i] = lp;
    if (logp_rec) logp_rec[i] = lp;
    if (mask_rec) mask_rec[i] = mk;
  }
}

// Normalizer.normalize for a block of rows (normalizer.py:107-110): out = (x - mean) / std, 128-bit when dim % 4 == 0
__global__ void obs_normalize_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                     const float* __restrict__ sd, long long rows, int dim, int ld_out, float* __restrict__ out,
                                     uint16_t* __restrict__ out16, uint32_t* slot) {
  const int q = ld_out / 4;                          // float4 slots per output row; slots >= dim / 4 are zero padding
  const long long n4 = rows * q;
  float vm = 0.f;                                    // max|out| of what this thread writes (f16x3: left in `slot`)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / q;
    const int c = (int)(i - r * q) * 4;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < dim) {
      const float4 v = ldg4(x + r * dim + c);
      const float4 m = ldg4(mean + c), s = ldg4(sd + c);
      o = make_float4(sub_rn(v.x, m.x) / s.x, sub_rn(v.y, m.y) / s.y, sub_rn(v.z, m.z) / s.z, sub_rn(v.w, m.w) / s.w);
    }
    stg4(out + 4 * i, o);
    vm = fmaxf(fmaxf(vm, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
    if (out16) { out16[4 * i] = to_bf16(o.x); out16[4 * i + 1] = to_bf16(o.y); out16[4 * i + 2] = to_bf16(o.z); out16[4 * i + 3] = to_bf16(o.w); }
  }
  amax_commit(vm, slot);
}

// wd0_pad[r, c] = c < dim ? W[r, c] : 0 : the discriminator's first-layer weight with 16-byte aligned rows, so that
// TMA can address it (disc_obs_dim = 114 floats = 456 bytes per row is not a legal tensor-map stride)
__global__ void pad_rows_kernel(const float* __restrict__ w, int rows, int dim, int ld, float* __restrict__ out,
                                uint16_t* __restrict__ out16) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ld) return;
  const int r = i / ld, c = i - r * ld;
  const float v = c < dim ? w[(size_t)r * dim + c] : 0.f;
  out[i] = v;
  if (out16) out16[i] = to_bf16(v);
}

__global__ void f32_to_bf16_flat_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = to_bf16(src[i]);
}

__global__ void diff_normalize_kernel(const float* __restrict__ dobs, const float* __restrict__ demo,
                                      const float* __restrict__ mean_abs, long long rows, int dim, int ld,
                                      float* __restrict__ out, uint16_t* __restrict__ out16) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ld) return;
  long long r = i / ld; int c = (int)(i - r * ld);
  const float v = c < dim ? sub_rn(demo[r * dim + c], dobs[r * dim + c]) / fmaxf(mean_abs[c], 1e-4f) : 0.f;
  out[i] = v;
  if (out16) out16[i] = to_bf16(v);
}

}  // namespace addk

using namespace addk;
typedef addk_update_ctx Ctx;
#define F(p) ((float*)(p))

// One chain's activations / gradients and its column-sum scratch.  The actor, the critic and the discriminator
// chains of an optimizer step each own one when they run on separate streams.
struct ChainWs {
  float *h1, *h2, *h3, *g1, *g2, *g3;
  float* colsum_work;
  float* colpart;       // [148 * 8, 1024] per-block column sums left behind by the f16x3 split pass (NULL: not available)
};
// wgrad() asks the conversion of its dY operand (inside gemm()) to leave column partial sums behind
static thread_local const float* g_colpart_for = nullptr;
static thread_local float* g_colpart_buf = nullptr;
static thread_local int g_colpart_rows = 0;
// ... or the dense layer that PRODUCED dY left them behind in its epilogue (gemm(..., colpart_out)): tensor and row count
static thread_local const float* g_colpart_epi_for = nullptr;
static thread_local int g_colpart_epi_rows = 0;

static int colsum(cudaStream_t st, const addk_update_ctx& c, const ChainWs& ws, const float* dY, int ld, int rows, int n,
                  float* out, const float* rw, const uint16_t* dY16 = nullptr, const uint32_t* planes_slot = nullptr) {
  using addk::COLSUM_MAX_N; using addk::COLSUM_CHUNKS;
  if (n > COLSUM_MAX_N) { addk_set_error("colsum: more than 1024 columns"); return ADDK_ERR_UNSUPPORTED; }
  float* work = ws.colsum_work;
  unsigned int* tickets = (unsigned int*)(work + (size_t)COLSUM_CHUNKS * COLSUM_MAX_N);
  const addk::PlanesArg none{nullptr, 0, nullptr};
  const int groups = (n + 127) / 128;
  int chunks = (148 * 8) / groups;                 // ~8 blocks per SM in total
  chunks = chunks > COLSUM_CHUNKS ? COLSUM_CHUNKS : (chunks < 8 ? 8 : chunks);
  if (chunks > rows) chunks = rows > 0 ? rows : 1;
  if (dY16 && planes_slot && (ld & 3) == 0)      // no fp32 copy (precision "f16x3"): read the fp16 planes
    addk::colsum_slabs_kernel<addk::PlanesPtr><<<dim3(groups, chunks), 256, 0, st>>>(
        nullptr, addk::PlanesArg{reinterpret_cast<const __half*>(dY16), c.arena_elems, planes_slot}, ld, rows, n, out, c.num_params,
        (int)c.split_k, rw, work, tickets);
  else if (dY16 && (ld & 3) == 0)      // the tensor has no fp32 copy (precision "bf16"): read its bf16 copy
    addk::colsum_slabs_kernel<uint16_t><<<dim3(groups, chunks), 256, 0, st>>>(dY16, none, ld, rows, n, out, c.num_params,
                                                                                     (int)c.split_k, rw, work, tickets);
  else
    addk::colsum_slabs_kernel<float><<<dim3(groups, chunks), 256, 0, st>>>(dY, none, ld, rows, n, out, c.num_params,
                                                                                  (int)c.split_k, rw, work, tickets);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}

// precision "bf16": the bf16 twin of an fp32 operand (NULL when the pointer is not one of the twinned tensors)
static thread_local const addk_update_ctx* g_twin_ctx = nullptr;

// precision "f16x3": every arena tensor has an fp16 hi/lo twin (arena16: hi plane, then the lo plane arena_elems later)
// and a max|x| word; a twin is (re)written by the first dense layer that reads the tensor after it was produced and
// reused by the later ones (an activation feeds the next layer AND its weight gradient, a gradient feeds the input
// gradient AND the weight gradient).  This table is host-side bookkeeping in ISSUE order; the kernels themselves are
// stream-ordered, so a twin may only be shared by calls on one stream -- tensors read by several chains (xn, dn, the
// parameters) are converted before the streams fork.  Slot 0 belongs to the flat parameter vector.
struct TwinEnt { const float* p; long long rows; int cols, ld; cudaStream_t st; bool valid, shared, amax_known; int slot; };
constexpr int TWIN_ENTRIES = 63, TWIN_SLOTS = 128;       // slots per entry point; slot 0 of the array = the parameters
static thread_local TwinEnt g_tw[TWIN_ENTRIES];
static thread_local int g_ntw = 0;
static thread_local int g_slot_base = 0;       // each entry point owns TWIN_SLOTS slots
static thread_local int g_next_slot = 0;       // slots [TWIN_ENTRIES, TWIN_SLOTS) are handed out once per call, pre-zeroed
// ReLU masks as bit planes (f16x3, optimizer step only): which arena tensors currently have a valid bit plane in
// arena_bits (written by the epilogue of the ReLU layer that produced them), in ISSUE order like the twin table
struct BitsEnt { const float* p; long long rows; int cols, ld; cudaStream_t st; bool valid; };
constexpr int BITS_ENTRIES = 16;
static thread_local BitsEnt g_bits[BITS_ENTRIES];
static thread_local int g_nbits = 0;
static thread_local bool g_want_bits = false;      // set by addk_update_minibatch (inference entry points do not need masks)
static thread_local bool g_prepped = false;        // the slots of this call were prepared in one launch (h3_params)
static void twin_reset(int entry_point) { g_ntw = 0; g_slot_base = TWIN_SLOTS * entry_point; g_next_slot = TWIN_ENTRIES; g_nbits = 0; g_want_bits = false; }
static uint32_t* slot_ptr(const addk_update_ctx& c, int slot) { return (uint32_t*)c.amax_slots + 2 * (1 + g_slot_base + slot); }
static uint32_t* twin_slot(const addk_update_ctx& c, int e) { return slot_ptr(c, g_tw[e].slot); }
static void twin_invalidate(const void* p, size_t bytes) {
  const char* b = (const char*)p;
  for (int i = 0; i < g_ntw; ++i) {
    const char* q = (const char*)g_tw[i].p;
    if (q >= b && q < b + (bytes ? bytes : 1)) { g_tw[i].valid = false; g_tw[i].amax_known = false; }
  }
  for (int i = 0; i < g_nbits; ++i) {
    const char* q = (const char*)g_bits[i].p;
    if (q >= b && q < b + (bytes ? bytes : 1)) g_bits[i].valid = false;
  }
}
// word pointer of the bit plane of arena tensor p (NULL: no bit planes / not an arena tensor / not on a 128-element boundary)
static uint32_t* bits_ptr(const addk_update_ctx& c, const float* p, int ld) {
  const float* a0 = (const float*)c.arena;
  if (!c.arena_bits || !a0 || !p || p < a0 || p >= a0 + c.arena_elems || ((p - a0) & 127) || (ld & 127)) return nullptr;
  return (uint32_t*)c.arena_bits + (p - a0) / 32;
}
struct H3Op { const void* hi; long long plane; uint32_t* amax; int ready; long long full_rows; int full_cols; };
// looks the operand up; on a miss the entry is created / refreshed and ready = 0 tells the library to convert first
static H3Op h3_operand(const addk_update_ctx& c, cudaStream_t st, const float* p, long long rows, int cols, int ld) {
  H3Op o{nullptr, 0, nullptr, 0, rows, cols};
  uint32_t* slots = (uint32_t*)c.amax_slots;
  if (!slots || !p) return o;
  const float* p0 = (const float*)c.params;
  if (c.params16 && p >= p0 && p < p0 + c.num_params) {
    o.hi = (const uint16_t*)c.params16 + (p - p0); o.plane = c.num_params; o.amax = slots; o.ready = 1;
    return o;
  }
  const float* a0 = (const float*)c.arena;
  if (!a0 || !c.arena16 || p < a0 || p >= a0 + c.arena_elems) return o;
  int e = -1;
  for (int i = 0; i < g_ntw; ++i) if (g_tw[i].p == p) { e = i; break; }
  if (e < 0) {
    if (g_ntw >= TWIN_ENTRIES) return o;
    e = g_ntw++;
    g_tw[e] = TwinEnt{p, 0, 0, 0, nullptr, false, false, false, e};
  }
  TwinEnt& t = g_tw[e];
  o.hi = (const uint16_t*)c.arena16 + (p - a0); o.plane = c.arena_elems; o.amax = twin_slot(c, e);
  // (a twin of [t.rows, t.cols] with the same pitch also serves a request for fewer rows / columns)
  if (t.valid && cols <= t.cols && t.ld == ld && rows <= t.rows && (t.st == st || t.shared)) { o.ready = 1; return o; }
  // the call about to be issued converts it; the max pass is skipped when the producing dense layer left max|x| behind
  if (t.amax_known && cols <= t.cols && t.ld == ld && rows <= t.rows && t.st == st) {
    o.ready = 2;
    rows = t.rows; cols = t.cols;       // the caller converts everything max|x| covers: later, wider requests are served too
    o.full_rows = rows; o.full_cols = cols;
  }
  t.rows = rows; t.cols = cols; t.ld = ld; t.st = st; t.valid = true; t.shared = false; t.amax_known = false;
  return o;
}
extern "C" int addk_f16x3_convert(void* stream, const float* x, long long rows, int cols, int ld, void* hi16, long long plane,
                                  uint32_t* amax_slot);
extern "C" int addk_f16x3_split(void* stream, const float* x, long long rows, int cols, int ld, void* hi16, long long plane,
                                uint32_t* amax_slot, float* colsum_partials, int* colsum_partial_rows);
extern "C" int addk_f16x3_prep(void* stream, uint32_t* slot, int keep_sticky_word);
extern "C" int addk_f16x3_prep_all(void* stream, uint32_t* slots, int n_sticky, int n_total);
extern "C" int addk_f16x3_repair(void* stream, const float* x, long long rows, int cols, int ld, void* hi16, long long plane,
                                 uint32_t* slot);
// convert now (used for tensors several streams read: must happen before the fork)
static int h3_prepare(const addk_update_ctx& c, cudaStream_t st, const float* p, long long rows, int cols, int ld) {
  H3Op o = h3_operand(c, st, p, rows, cols, ld);
  if (!o.hi) { addk_set_error("f16x3: tensor has no twin"); return ADDK_ERR_ARG; }
  for (int i = 0; i < g_ntw; ++i) if (g_tw[i].p == p) g_tw[i].shared = true;     // converted before the fork: any stream may read it
  if (o.ready == 1) return ADDK_OK;
  if (o.ready == 2)                     // the producing kernel left max|x| behind (amax_hook): the split pass alone
    return addk_f16x3_split(st, p, o.full_rows, o.full_cols, ld, const_cast<void*>(o.hi), o.plane, o.amax, nullptr, nullptr);
  return addk_f16x3_convert(st, p, rows, cols, ld, const_cast<void*>(o.hi), o.plane, o.amax);
}
// An elementwise kernel is about to write the arena tensor p ([rows, cols], pitch ld) on stream st and will leave max|p|
// in the returned slot (a fresh one, zeroed by h3_params); NULL = not in f16x3 mode / no slot left (the conversion
// then runs its own max pass).  Call AFTER twin16(p), which forgets the old twin.
static uint32_t* amax_hook(cudaStream_t st, const void* p, long long rows, int cols, int ld) {
  const addk_update_ctx* c = g_twin_ctx;
  if (!c || c->precision != 4 || !c->amax_slots || !p || g_next_slot >= TWIN_SLOTS) return nullptr;
  if (!addk_switches().h3_amax_hooks) return nullptr;   // ADDK_H3_AMAX_HOOKS=0: A/B switch (separate max passes)
  const float* f = (const float*)p;
  const float* a0 = (const float*)c->arena;
  if (!a0 || f < a0 || f >= a0 + c->arena_elems) return nullptr;
  int e = -1;
  for (int i = 0; i < g_ntw; ++i) if (g_tw[i].p == f) { e = i; break; }
  if (e < 0) {
    if (g_ntw >= TWIN_ENTRIES) return nullptr;
    e = g_ntw++;
    g_tw[e] = TwinEnt{f, 0, 0, 0, nullptr, false, false, false, e};
  }
  TwinEnt& t = g_tw[e];
  t.slot = g_next_slot++;
  t.rows = rows; t.cols = cols; t.ld = ld; t.st = st; t.valid = false; t.shared = false; t.amax_known = true;
  return twin_slot(*c, e);
}
// the fp16 planes of the whole flat parameter vector (one scale: slot 0)
static int h3_params(const addk_update_ctx& c, cudaStream_t st, int entry_point) {
  if (c.precision != 4) return ADDK_OK;
  twin_reset(entry_point);
  if (!c.params16 || !c.amax_slots) { addk_set_error("f16x3: the context has no parameter twin / max|x| slots"); return ADDK_ERR_ARG; }
  // every slot of this entry point in ONE launch: the per-tensor slots keep their sticky scale word (W <- the scale in
  // force, max <- 0: what a single-thread launch in front of every planes-writing layer used to do, twelve per optimizer
  // step on the chains' critical paths), the once-per-call slots are cleared (a dense layer that leaves max|C| behind takes
  // a fresh one).  A slot written twice in one call (u1 of the discriminator chain) accumulates the larger max|x|: its
  // scale then fits both, and the readers derive the scale from the two words as before.
  { const int rcp = addk_f16x3_prep_all(st, slot_ptr(c, 0), TWIN_ENTRIES, TWIN_SLOTS); if (rcp != ADDK_OK) return rcp; }
  g_prepped = true;
  if (c.params16_current) return ADDK_OK;      // (addk_params_refresh ran after the last parameter change)
  return addk_f16x3_convert(st, (const float*)c.params, 1, (int)c.num_params, (int)c.num_params, c.params16, c.num_params,
                            (uint32_t*)c.amax_slots);
}

static uint16_t* twin16(const void* p) {
  const addk_update_ctx* c = g_twin_ctx;
  if (!c || !p) return nullptr;
  if (c->precision == 4) { twin_invalidate(p, 0); return nullptr; }   // an elementwise producer is about to rewrite p
  if (c->precision != 3) return nullptr;
  const float* f = (const float*)p;
  const float* a0 = (const float*)c->arena;
  if (a0 && f >= a0 && f < a0 + c->arena_elems) return (uint16_t*)c->arena16 + (f - a0);
  const float* p0 = (const float*)c->params;
  if (c->params16 && f >= p0 && f < p0 + c->num_params) return (uint16_t*)c->params16 + (f - p0);
  return nullptr;
}

enum { F16_DROP_C = 1, F16_MASK = 2 };
static bool drop16_enabled(int prec) { return prec == 3 ? addk_switches().bf16_drop_f32 != 0 : addk_switches().h3_planes_only != 0; }
// would a [rows, cols] layer output be 16-bit only in this mode?  (same rule as gemm(): bf16 / f16x3 mode, persistent kernel)
static bool is_16only(const addk_update_ctx& c, long long rows, int cols) {
  if ((c.precision != 3 && c.precision != 4) || !drop16_enabled((int)c.precision) || !c.arena16) return false;
  if (c.precision == 4) return cols > 128 && (cols & 7) == 0 && rows >= 64 * c.split_k && c.amax_slots;
  // every consumer must be able to run on the 16-bit copy: the split-K weight gradient (contraction over the rows) needs
  // at least one 64-row k-block per slab, otherwise it falls back to the fp32 operands
  if (rows < 64 * c.split_k) return false;
  addk_gemm_args a{};
  a.M = (int)rows; a.N = cols; a.split_k = 1;
  return addk_gemm_is_persistent(&a, 3) != 0;
}

static int gemm(cudaStream_t st, int prec, const float* A, int lda, int ta, const float* B, int ldb, int tb, float* C,
                int ldc, int M, int N, int K, const float* bias = nullptr, int relu = 0, const float* mask = nullptr,
                int ld_mask = 0, int split = 1, const float* nmean = nullptr, const float* nstd = nullptr,
                long long slab_stride = 0, int flags16 = 0, float* colpart_out = nullptr) {
  // flags16 (precision "bf16"): F16_DROP_C -- C is only ever read by dense layers / as a ReLU mask / by colsum: do not
  // write its fp32 copy when the persistent kernel runs; F16_MASK -- the mask source was produced that way: read its bf16 copy
  addk_gemm_args a;
  a.relu_mask_src16 = nullptr; a.no_f32 = 0;
  bool planes_only = false;
  a.A16 = prec == 3 ? twin16(A) : nullptr; a.B16 = prec == 3 ? twin16(B) : nullptr; a.C16 = prec == 3 ? twin16(C) : nullptr;
  a.a16_plane = a.b16_plane = a.c16_plane = 0; a.a_amax = a.b_amax = a.c_amax = nullptr; a.a16_ready = a.b16_ready = 0;
  int c_ent = -1;
  a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K;
  a.bias = bias; a.a_mean = nmean; a.a_std = nstd; a.relu_mask_src = mask; a.ld_mask = ld_mask;
  a.trans_a = ta; a.trans_b = tb; a.relu = relu; a.split_k = split; a.accumulate = 0; a.slab_stride = slab_stride;
  if (prec == 4 && g_twin_ctx) {
    if (addk_gemm_h3_usable(a)) {
      const H3Op oa = h3_operand(*g_twin_ctx, st, A, ta ? K : M, ta ? M : K, lda);
      const H3Op ob = h3_operand(*g_twin_ctx, st, B, tb ? N : K, tb ? K : N, ldb);
      if (oa.hi && ob.hi) {
        a.A16 = oa.hi; a.a16_plane = oa.plane; a.a_amax = oa.amax; a.a16_ready = oa.ready;
        a.B16 = ob.hi; a.b16_plane = ob.plane; a.b_amax = ob.amax; a.b16_ready = ob.ready;
        // max|x| already known (left by the producer): split the WHOLE region it covers here, not just this call's view
        if (oa.ready == 2) {
          const bool want = g_colpart_buf && g_colpart_for == A && oa.full_rows == (ta ? K : M) && oa.full_cols == (ta ? M : K);
          const int rc = addk_f16x3_split(st, A, oa.full_rows, oa.full_cols, lda, const_cast<void*>(oa.hi), oa.plane, oa.amax,
                                          want ? g_colpart_buf : nullptr, want ? &g_colpart_rows : nullptr);
          if (rc != ADDK_OK) return rc;
          a.a16_ready = 1;
        }
        if (ob.ready == 2) {
          const int rc = addk_f16x3_split(st, B, ob.full_rows, ob.full_cols, ldb, const_cast<void*>(ob.hi), ob.plane, ob.amax,
                                          nullptr, nullptr);
          if (rc != ADDK_OK) return rc;
          a.b16_ready = 1;
        }
      } else {                      // the call runs in tf32x3 and converts nothing: forget what h3_operand assumed
        if (oa.hi && oa.ready != 1) twin_invalidate(A, 0);
        if (ob.hi && ob.ready != 1) twin_invalidate(B, 0);
      }
    }
    twin_invalidate(C, (size_t)(split > 1 ? 1 : M) * ldc * sizeof(float));
    if (a.A16 && split == 1) {      // the f16x3 kernel runs: let its epilogue leave max|C| in C's slot for C's first reader
      const addk_update_ctx& c = *g_twin_ctx;
      const float* a0 = (const float*)c.arena;
      if (a0 && C >= a0 && C < a0 + c.arena_elems) {
        int e = -1;
        for (int i = 0; i < g_ntw; ++i) if (g_tw[i].p == C) { e = i; break; }
        if (e < 0 && g_ntw < TWIN_ENTRIES) { e = g_ntw++; g_tw[e] = TwinEnt{C, 0, 0, 0, nullptr, false, false, false, e}; }
        if (e >= 0) {
          // Letting the epilogue write C's planes too (sticky scale word, prep -> layer -> repair) is implemented and
          // tested but off: measured at 4096 envs it removes 0.18 ms of split passes per optimizer step and adds 0.25 ms
          // to the dense layers (the epilogue's 8-byte stores are far from the split kernel's 6 TB/s).
          // fused = 2: only behind a long contraction (K >= 1024), where the worker warps wait for the tensor core anyway
          const int fused_sw = addk_switches().h3_fused_planes;
          // planes only: the layer's output is read by dense layers / as a ReLU mask / by the bias column sums alone, so it
          // never exists in fp32 -- the epilogue writes the two fp16 planes with the scale of the previous optimizer step
          // (sticky word) and a second launch repairs them in the rare case that scale does not fit (gemm_tc.cu: repair)
          planes_only = (flags16 & F16_DROP_C) && is_16only(c, M, N);
          const bool fused = planes_only || (fused_sw && N > 128 && (ldc & 7) == 0 && (fused_sw != 2 || K >= 1024));
          bool ready = false;
          if (!fused && g_next_slot < TWIN_SLOTS) { g_tw[e].slot = g_next_slot++; ready = true; }     // zeroed by h3_params
          else { g_tw[e].slot = e; }
          uint32_t* slot = twin_slot(c, e);
          if (ready || g_prepped || addk_f16x3_prep(st, slot, fused ? 1 : 0) == ADDK_OK) {     // max <- 0, W <- scale in force | 0
            TwinEnt& t = g_tw[e];
            t.rows = M; t.cols = N; t.ld = ldc; t.st = st; t.valid = false; t.shared = false; t.amax_known = true;
            a.c_amax = slot;
            if (fused) {     // the persistent kernel also writes C's planes in its epilogue
              a.C16 = (uint16_t*)c.arena16 + (C - a0); a.c16_plane = c.arena_elems;
              c_ent = e;
              if (planes_only) a.no_f32 = 1;
            }
          }
        }
      }
    }
  }
  if (prec == 3 && flags16 && a.A16 && a.B16 && a.C16 && g_twin_ctx && is_16only(*g_twin_ctx, M, N)) {
    if (flags16 & F16_DROP_C) a.no_f32 = 1;
    if ((flags16 & F16_MASK) && mask) { a.relu_mask_src16 = twin16(mask); if (a.relu_mask_src16) a.relu_mask_src = nullptr; }
  }
  if (prec == 4 && (flags16 & F16_DROP_C) && !planes_only && g_twin_ctx && is_16only(*g_twin_ctx, M, N)) {
    // the rule said "planes only" but the call cannot deliver them (no twin / tf32 fallback): consumers would read stale planes
    addk_set_error("f16x3: a planes-only layer fell off the fp16 path");
    return ADDK_ERR_UNSUPPORTED;
  }
  if (prec == 4 && (flags16 & F16_MASK) && mask && g_twin_ctx && is_16only(*g_twin_ctx, M, N)) {
    // the mask source (same shape as this output) exists only as planes: sign test on its hi plane
    const addk_update_ctx& c = *g_twin_ctx;
    a.relu_mask_src16 = (const uint16_t*)c.arena16 + (mask - (const float*)c.arena);
    a.relu_mask_src = nullptr;
  }
  a.relu_bits_out = nullptr; a.relu_bits_in = nullptr; a.ld_bits = 0;
  // the bias gradient of a gradient tensor that only exists as planes / its 16-bit copy: column sums per 32-row block
  // from the epilogue of the layer that produces it (wgrad() reduces them), instead of a second pass over the tensor
  a.colsum_partials = nullptr;
  if (colpart_out || C == g_colpart_epi_for) g_colpart_epi_for = nullptr;      // (the buffer / the tensor is about to be rewritten)
  if (colpart_out && (prec == 4 || prec == 3) && g_twin_ctx && a.A16 && a.B16 && split == 1 && (N & 63) == 0 && (ldc & 3) == 0 &&
      addk_switches().h3_colpart && addk_gemm_is_persistent(&a, prec) && (prec == 4 || (a.C16 && M > 128)) &&
      (long long)((M + 31) / 32) <= 148 * 8) {
    a.colsum_partials = colpart_out;
  }
  int bits_ent = -1;
  if ((prec == 4 || prec == 3) && g_twin_ctx && g_want_bits && a.A16 && a.B16 && split == 1 && (N & 127) == 0 && addk_gemm_is_persistent(&a, prec) &&
      (prec == 4 || (a.C16 && M > 128))) {      // (bf16: the calls gemm_bf16() routes to the persistent kernel)
    const addk_update_ctx& c = *g_twin_ctx;
    if (mask) {                  // the mask source's ReLU bits, if the layer that produced it left them behind (same stream)
      for (int i = 0; i < g_nbits; ++i) {
        const BitsEnt& b = g_bits[i];
        if (b.p == mask && b.valid && b.st == st && b.ld == ld_mask && N <= b.cols && M <= b.rows) {
          a.relu_bits_in = bits_ptr(c, mask, ld_mask); a.ld_bits = ld_mask / 32;
          break;
        }
      }
    } else if (relu) {           // a ReLU layer of the optimizer step: its output is (also) the mask of the backward pass
      uint32_t* bo = bits_ptr(c, C, ldc);
      if (bo) {
        for (int i = 0; i < g_nbits; ++i) if (g_bits[i].p == C) { bits_ent = i; break; }
        if (bits_ent < 0 && g_nbits < BITS_ENTRIES) bits_ent = g_nbits++;
        if (bits_ent >= 0) {
          g_bits[bits_ent] = BitsEnt{C, M, N, ldc, st, false};
          a.relu_bits_out = bo; a.ld_bits = ldc / 32;
        }
      }
    }
  }
  int rc = prec == 0 ? addk::sgemm_launch(st, a) : addk_gemm_tc(st, a, prec);
  if (rc != ADDK_OK) return rc;
  ADDK_CHECK_LAUNCH();
  if (bits_ent >= 0) g_bits[bits_ent].valid = true;
  if (a.colsum_partials) { g_colpart_epi_for = C; g_colpart_epi_rows = (M + 31) / 32; }
  if (c_ent >= 0) {     // planes written by the epilogue: rewritten only if the sticky scale did not fit max|C|
    if (!planes_only) {   // (planes-only layers repair themselves: second launch inside addk_gemm)
      rc = addk_f16x3_repair(st, C, M, N, ldc, a.C16, a.c16_plane, a.c_amax);
      if (rc != ADDK_OK) return rc;
    }
    g_tw[c_ent].valid = true;
  }
  return ADDK_OK;
}
#define TRY(x) do { int rc__ = (x); if (rc__ != ADDK_OK) return rc__; } while (0)

// weight gradient dW[N_out, K_in] = dY^T X and bias gradient db = 1^T dY, as split-K slabs
static int wgrad(cudaStream_t st, const Ctx& c, const ChainWs& ws, const float* dY, int ldy, const float* X, int ldx,
                 int rows, int n_out, int k_in, long long o_w, long long o_b, int slab0, bool dy_16only = false) {
  const int S = (int)c.split_k;
  const long long P = c.num_params;
  if (n_out == 1 && ldy == 1) {
    TRY(colsum(st, c, ws, X, ldx, rows, k_in, F(c.slabs) + (size_t)slab0 * P + o_w, dY));
  } else {
    // f16x3: if this call is the one that converts dY, its split pass also leaves the column sums of dY (the bias gradient)
    const int fuse = addk_switches().h3_colpart;          // ADDK_H3_COLPART=0: A/B switch (separate column-sum kernels)
    g_colpart_rows = 0;
    const bool have_epi = ws.colpart && g_colpart_epi_for == dY;      // the layer that produced dY left the partials already
    g_colpart_for = (o_b >= 0 && fuse && !have_epi) ? dY : nullptr;
    g_colpart_buf = (o_b >= 0 && fuse && !have_epi) ? ws.colpart : nullptr;
    const int rc = gemm(st, (int)c.precision, dY, ldy, 1, X, ldx, 0, F(c.slabs) + (size_t)slab0 * P + o_w, k_in, n_out, k_in, rows,
                        nullptr, 0, nullptr, 0, S, nullptr, nullptr, P);
    g_colpart_for = nullptr; g_colpart_buf = nullptr;
    if (rc != ADDK_OK) return rc;
    if (o_b >= 0 && g_colpart_rows > 0) {
      // (scratch: the chain's colsum work area -- the two kernels never overlap inside one chain; tickets live behind the
      //  64 chunk rows of colsum_slabs_kernel, a second set of 32 counters)
      float* cwork = ws.colsum_work;
      unsigned int* ctick = (unsigned int*)(cwork + (size_t)COLSUM_CHUNKS * COLSUM_MAX_N) + 32;
      colsum_parts_kernel<<<dim3((n_out + 31) / 32, CP_CHUNKS), 256, 0, st>>>(ws.colpart, g_colpart_rows, n_out,
                                                                               F(c.slabs) + (size_t)slab0 * P + o_b, P, S, cwork, ctick);
      ADDK_CHECK_LAUNCH();
      g_colpart_rows = 0;
      return ADDK_OK;
    }
  }
  if (o_b >= 0 && ws.colpart && g_colpart_epi_for == dY && g_colpart_epi_rows == (rows + 31) / 32) {
    float* cwork = ws.colsum_work;
    unsigned int* ctick = (unsigned int*)(cwork + (size_t)COLSUM_CHUNKS * COLSUM_MAX_N) + 32;
    colsum_parts_kernel<<<dim3((n_out + 31) / 32, CP_CHUNKS), 256, 0, st>>>(ws.colpart, g_colpart_epi_rows, n_out,
                                                                             F(c.slabs) + (size_t)slab0 * P + o_b, P, S, cwork, ctick);
    ADDK_CHECK_LAUNCH();
    g_colpart_epi_for = nullptr;
    return ADDK_OK;
  }
  if (o_b >= 0) {
    const bool only16 = dy_16only && is_16only(c, rows, n_out);
    const uint16_t* dY16 = nullptr;
    const uint32_t* pslot = nullptr;
    if (only16 && c.precision == 3) dY16 = twin16(dY);
    if (only16 && c.precision == 4) {     // the fp16 planes of dY and the slot that holds their scale
      for (int i = 0; i < g_ntw; ++i) if (g_tw[i].p == dY && g_tw[i].valid) { pslot = twin_slot(c, i); break; }
      if (!pslot) { addk_set_error("f16x3: planes-only gradient without a valid twin"); return ADDK_ERR_UNSUPPORTED; }
      dY16 = (const uint16_t*)c.arena16 + (dY - (const float*)c.arena);
    }
    TRY(colsum(st, c, ws, dY, ldy, rows, n_out, F(c.slabs) + (size_t)slab0 * P + o_b, nullptr, dY16, pslot));
  }
  return ADDK_OK;
}

static int head1_forward(cudaStream_t st, const float* X, int ld, long long rows, int K, const float* w, const float* b,
                         float* out) {
  rowdot_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(X, ld, rows, K, w, b, out);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}
static int head1_dgrad(cudaStream_t st, const float* d, const float* w, const float* h, long long rows, int K, float* g) {
  const long long n4 = rows * K / 4;
  long long bl = (n4 + 255) / 256; if (bl > 148 * 16) bl = 148 * 16;
  uint16_t* g16 = twin16(g);
  outer_mask_kernel<<<(unsigned)bl, 256, 0, st>>>(d, w, h, rows, K, g, g16, amax_hook(st, g, rows, K, K));
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}

// 3-hidden-layer trunk forward: X[rows,in] -> h1,h2,h3
static int trunk_forward(cudaStream_t st, const Ctx& c, const ChainWs& ws, const float* X, int ldx, int in_dim, int rows, long long o_w0,
                         long long o_b0, long long o_w1, long long o_b1, long long o_w2, long long o_b2,
                         const float* nmean = nullptr, const float* nstd = nullptr, float* w0_pad = nullptr) {
  const float* P = F(c.params);
  const int H1 = (int)c.hid_a1, H2 = (int)c.hid_a2, H3 = (int)c.hid_a3, pr = (int)c.precision;
  const int OL = (int)c.obs_ld;
  if (nmean && pr != 0 && (in_dim & 3) == 0 && ldx == in_dim && rows <= c.mb_rows + 1) {
    // tensor-core modes: normalise into the minibatch scratch first (TMA cannot apply it on load), then the TC tile
    const long long n4 = (long long)rows * OL / 4;
    int bl = (int)((n4 + 255) / 256); if (bl > 148 * 8) bl = 148 * 8;
    uint16_t* const xn16 = twin16(c.xn);
    obs_normalize_kernel<<<bl, 256, 0, st>>>(X, nmean, nstd, rows, in_dim, OL, F(c.xn), xn16, amax_hook(st, c.xn, rows, OL, OL));
    ADDK_CHECK_LAUNCH();
    X = F(c.xn); ldx = OL; nmean = nullptr; nstd = nullptr;
  }
  // First-layer weight [H1, in_dim]: a row pitch of 264 elements is 528 bytes in the 16-bit twins -- legal for TMA (16)
  // but every other row starts in the middle of a 32-byte sector, and the layer's main loop then waits for its loads
  // (measured 60 us at K = 264 against 39 / 44 us at K = 256 / 288).  The tensor-core modes therefore read a copy
  // with rows padded to obs_ld (a multiple of 16), like the discriminator's wd0_pad; xn has the same pitch.
  const float* W0 = P + o_w0;
  int ldw = in_dim;
  if (w0_pad && !nmean && pr != 0 && X == F(c.xn) && OL > in_dim) {
    if (!c.params16_current) {
      pad_rows_kernel<<<(H1 * OL + 255) / 256, 256, 0, st>>>(P + o_w0, H1, in_dim, OL, w0_pad, twin16(w0_pad));
      ADDK_CHECK_LAUNCH();
    }
    W0 = w0_pad; ldw = OL;
  }
  // (h1, h2 are read by dense layers and as ReLU masks only: 16-bit only in bf16 mode)
  TRY(gemm(st, nmean ? 0 : pr, X, ldx, 0, W0, ldw, 1, ws.h1, H1, rows, H1, in_dim, P + o_b0, 1, nullptr, 0, 1,
           nmean, nstd, 0, F16_DROP_C));
  TRY(gemm(st, pr, ws.h1, H1, 0, P + o_w1, H1, 1, ws.h2, H2, rows, H2, H1, P + o_b1, 1, nullptr, 0, 1, nullptr, nullptr, 0, F16_DROP_C));
  TRY(gemm(st, pr, ws.h2, H2, 0, P + o_w2, H2, 1, ws.h3, H3, rows, H3, H2, P + o_b2, 1));
  return ADDK_OK;
}

// trunk backward given g3 = dL/dh3 (already masked by h3 > 0)
static int trunk_backward(cudaStream_t st, const Ctx& c, const ChainWs& ws, const float* X, int ldx, int in_dim, int rows, long long o_w0,
                          long long o_b0, long long o_w1, long long o_b1, long long o_w2, long long o_b2) {
  const float* P = F(c.params);
  const int H1 = (int)c.hid_a1, H2 = (int)c.hid_a2, H3 = (int)c.hid_a3, pr = (int)c.precision;
  TRY(wgrad(st, c, ws, ws.g3, H3, ws.h2, H2, rows, H3, H2, o_w2, o_b2, 0));
  // (g2, g1: dense-layer operands + bias column sums only; their masks h2, h1 have no fp32 copy in bf16 mode)
  TRY(gemm(st, pr, ws.g3, H3, 0, P + o_w2, H2, 0, ws.g2, H2, rows, H2, H3, nullptr, 0, ws.h2, H2, 1, nullptr, nullptr, 0, F16_DROP_C | F16_MASK, ws.colpart));
  TRY(wgrad(st, c, ws, ws.g2, H2, ws.h1, H1, rows, H2, H1, o_w1, o_b1, 0, true));
  TRY(gemm(st, pr, ws.g2, H2, 0, P + o_w1, H1, 0, ws.g1, H1, rows, H1, H2, nullptr, 0, ws.h1, H1, 1, nullptr, nullptr, 0, F16_DROP_C | F16_MASK, ws.colpart));
  TRY(wgrad(st, c, ws, ws.g1, H1, X, ldx, rows, H1, in_dim, o_w0, o_b0, 0, true));
  return ADDK_OK;
}

static ChainWs main_ws(const Ctx& c) {
  return ChainWs{F(c.h1), F(c.h2), F(c.h3), F(c.g1), F(c.g2), F(c.g3), F(c.colsum_work), F(c.colpart_a)};
}

// Two helper streams (+ fork / join events) per device for the critic and the discriminator chains, created on first use.
struct AuxStreams {
  int device = -1;
  cudaStream_t s[2] = {nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
};
static int aux_streams(AuxStreams** out) {
  static AuxStreams a;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { addk_set_error("cudaGetDevice failed"); return ADDK_ERR_LAUNCH; }
  if (a.device != dev) {
    if (a.device >= 0) {   // the process moved to another device: drop the old helpers
      for (int i = 0; i < 2; ++i) { cudaStreamDestroy(a.s[i]); cudaEventDestroy(a.join[i]); }
      cudaEventDestroy(a.fork);
    }
    bool ok = cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i)
      ok = cudaStreamCreateWithFlags(&a.s[i], cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&a.join[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { a.device = -1; addk_set_error("could not create the helper streams of the update"); return ADDK_ERR_LAUNCH; }
    a.device = dev;
  }
  *out = &a;
  return ADDK_OK;
}

extern "C" int addk_clip_grad_norm(void* stream, float* grads, long long n, double max_norm, double pre_scale,
                                   double* sumsq_work, float* coef_out) {
  if (!grads || !sumsq_work || !coef_out || n <= 0 || !(max_norm > 0.0)) return ADDK_ERR_ARG;
  if (reinterpret_cast<uintptr_t>(grads) & 15) { addk_set_error("clip_grad_norm: the flat gradient must be 16-byte aligned"); return ADDK_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(sumsq_work, 0, sizeof(double), st);
  sumsq_kernel<<<148 * 4, 256, 0, st>>>(grads, n, sumsq_work);
  ADDK_CHECK_LAUNCH();
  clip_coef_kernel<<<1, 1, 0, st>>>(sumsq_work, (float)pre_scale, (float)max_norm, coef_out);
  ADDK_CHECK_LAUNCH();
  scale_by_coef_kernel<<<148 * 4, 256, 0, st>>>(grads, n, coef_out);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}

extern "C" int addk_update_ctx_size(void) { return (int)sizeof(Ctx); }

extern "C" int addk_update_ctx_init(void* ctx_host, const void* const* ptrs, int n_ptrs, const int64_t* ints,
                                    int n_ints, const double* floats, int n_floats) {
  if (!ctx_host || !ptrs || !ints || !floats) return ADDK_ERR_ARG;
  Ctx* c = (Ctx*)ctx_host;
  int ip = 0, ii = 0, id = 0;
#define ADDK_PTR(n) if (ip >= n_ptrs) return ADDK_ERR_ARG; c->n = (void*)ptrs[ip++];
#define ADDK_INT(n) if (ii >= n_ints) return ADDK_ERR_ARG; c->n = ints[ii++];
#define ADDK_F64(n) if (id >= n_floats) return ADDK_ERR_ARG; c->n = floats[id++];
#include "ctx_fields.h"
  if (ip != n_ptrs || ii != n_ints || id != n_floats) return ADDK_ERR_ARG;
  if (c->act_dim > 32 || c->split_k < 1 || c->split_k > 16) return ADDK_ERR_ARG;
  return ADDK_OK;
}

extern "C" int addk_update_minibatch(void* stream, void* ctx_host, const long long* idx, int step_index,
                                     int do_optim) {
  if (!ctx_host || !idx || step_index < 0) return ADDK_ERR_ARG;
  const Ctx& c = *(const Ctx*)ctx_host;
  if (c.params16_current) { addk_set_error("update_minibatch: a context with params16_current = 1 is for inference only"); return ADDK_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const int M = (int)c.mb_rows, R = M + 1, OD = (int)c.obs_dim, OL = (int)c.obs_ld, AD = (int)c.act_dim, AL = (int)c.act_ld;
  const int DD = (int)c.disc_dim, DL = (int)c.disc_ld, pr = (int)c.precision, S = (int)c.split_k;
  const int H1 = (int)c.hid_a1, H3 = (int)c.hid_a3, E1 = (int)c.hid_d1, E2 = (int)c.hid_d2;
  const long long P = c.num_params;
  const float* W = F(c.params);
  double* stats = (double*)c.stats;
  int* cnt = (int*)c.cnt;
  (void)H1;
  cudaMemsetAsync(stats, 0, ST_COUNT * sizeof(double), st);
  cudaMemsetAsync(cnt, 0, sizeof(int), st);
  g_twin_ctx = &c;
  if (c.precision == 3 && c.params16 && !c.params16_current) {
    f32_to_bf16_flat_kernel<<<148 * 4, 256, 0, st>>>(F(c.params), (uint16_t*)c.params16, c.num_params);
    ADDK_CHECK_LAUNCH();
  }
  TRY(h3_params(c, st, 0));
  g_nbits = 0;      // (bf16 mode does not go through twin_reset)
  g_want_bits = (c.precision == 4 || c.precision == 3) && c.arena_bits != nullptr && addk_switches().h3_relu_bits != 0;

  uint16_t* const xn16 = twin16(c.xn);
  uint16_t* const dn16 = twin16(c.dn);
  gather_minibatch_kernel<<<(M + 7) / 8, 256, 0, st>>>(
      idx, M, OD, OL, AD, AL, DD, DL, F(c.buf_obs), F(c.buf_action), F(c.buf_a_logp), F(c.buf_adv), F(c.buf_tar_val),
      F(c.buf_mask), F(c.buf_disc_obs), F(c.buf_disc_demo), F(c.obs_mean), F(c.obs_std), F(c.a_mean), F(c.a_std),
      F(c.disc_mean_abs), F(c.xn), F(c.an), F(c.old_logp), F(c.adv), F(c.tar), F(c.mask), F(c.dn), cnt, xn16, dn16,
      amax_hook(st, c.xn, M, OL, OL), amax_hook(st, c.dn, R, DL, DL));      // (row M of dn is the constant zero row)
  ADDK_CHECK_LAUNCH();

  if (pr == 4) {      // twins of the inputs the chains share, before the streams fork
    TRY(h3_prepare(c, st, F(c.xn), M, OD, OL));
    TRY(h3_prepare(c, st, F(c.dn), R, DL, DL));
  }

  // The three chains share only read-only inputs (xn, dn, the parameters) and write disjoint slab segments and
  // statistics slots.  With n_streams == 3 each has its own workspace and stream: the tail wave of one chain's dense
  // layer (3.46 waves of tiles at M = 16384) is filled by the other chains' tiles.
  const bool multi = c.n_streams == 3 && c.c_h1 && c.d_e1 && c.colsum_work_c && c.colsum_work_d;
  cudaStream_t sa = st, sc = st, sd = st;
  AuxStreams* aux = nullptr;
  if (multi) {
    TRY(aux_streams(&aux));
    sc = aux->s[0]; sd = aux->s[1];
    cudaEventRecord(aux->fork, st);
    cudaStreamWaitEvent(sc, aux->fork, 0);
    cudaStreamWaitEvent(sd, aux->fork, 0);
  }
  const ChainWs wa = main_ws(c);
  const ChainWs wc = multi ? ChainWs{F(c.c_h1), F(c.c_h2), F(c.c_h3), F(c.c_g1), F(c.c_g2), F(c.c_g3), F(c.colsum_work_c), F(c.colpart_c)} : wa;
  const ChainWs wd = multi ? ChainWs{F(c.d_e1), nullptr, F(c.d_e2), F(c.d_dv1), F(c.d_du2), F(c.d_dh2), F(c.colsum_work_d), F(c.colpart_d)} : wa;
  float* pred_d = multi ? F(c.d_pred) : F(c.pred);
  float* dpred_d = multi ? F(c.d_dpred) : F(c.dpred);

  // ---------------- actor (stream sa) ----------------
  TRY(trunk_forward(sa, c, wa, F(c.xn), OL, OD, M, c.o_a_w0, c.o_a_b0, c.o_a_w1, c.o_a_b1, c.o_a_w2, c.o_a_b2, nullptr, nullptr, F(c.wa0_pad)));
  TRY(gemm(sa, pr, wa.h3, H3, 0, W + c.o_a_wm, H3, 1, F(c.mean), AL, M, AD, H3, W + c.o_a_bm, 0));
  uint16_t* const dmean16 = twin16(c.dmean);
  actor_loss_kernel<<<(M + 7) / 8, 256, 0, sa>>>(F(c.mean), F(c.an), F(c.logstd), F(c.old_logp), F(c.adv), F(c.mask), M,
                                                 AD, AL, (float)c.ppo_clip_ratio, (float)c.action_bound_weight, cnt,
                                                 F(c.dmean), stats, dmean16, amax_hook(sa, c.dmean, M, AL, AL));
  ADDK_CHECK_LAUNCH();
  TRY(wgrad(sa, c, wa, F(c.dmean), AL, wa.h3, H3, M, AD, H3, c.o_a_wm, c.o_a_bm, 0));
  TRY(gemm(sa, pr, F(c.dmean), AL, 0, W + c.o_a_wm, H3, 0, wa.g3, H3, M, H3, AD, nullptr, 0, wa.h3, H3));
  TRY(trunk_backward(sa, c, wa, F(c.xn), OL, OD, M, c.o_a_w0, c.o_a_b0, c.o_a_w1, c.o_a_b1, c.o_a_w2, c.o_a_b2));

  // ---------------- critic (stream sc) ----------------
  TRY(trunk_forward(sc, c, wc, F(c.xn), OL, OD, M, c.o_c_w0, c.o_c_b0, c.o_c_w1, c.o_c_b1, c.o_c_w2, c.o_c_b2, nullptr, nullptr, F(c.wc0_pad)));
  TRY(head1_forward(sc, wc.h3, H3, M, H3, W + c.o_c_wo, W + c.o_c_bo, F(c.pred)));
  critic_loss_kernel<<<(M + 255) / 256, 256, 0, sc>>>(F(c.pred), F(c.tar), M, (float)c.critic_loss_weight, F(c.dpred),
                                                      stats);
  ADDK_CHECK_LAUNCH();
  TRY(wgrad(sc, c, wc, F(c.dpred), 1, wc.h3, H3, M, 1, H3, c.o_c_wo, c.o_c_bo, 0));
  TRY(head1_dgrad(sc, F(c.dpred), W + c.o_c_wo, wc.h3, M, H3, wc.g3));
  TRY(trunk_backward(sc, c, wc, F(c.xn), OL, OD, M, c.o_c_w0, c.o_c_b0, c.o_c_w1, c.o_c_b1, c.o_c_w2, c.o_c_b2));

  // ---------------- discriminator (stream sd; R = M + 1 rows) ----------------
  float *e1 = wd.h1, *e2 = wd.h3, *dh2 = wd.g3, *dv1 = wd.g1, *du2 = wd.g2;
  const float* Wd0 = F(c.wd0_pad);
  pad_rows_kernel<<<(E1 * DL + 255) / 256, 256, 0, sd>>>(W + c.o_d_w0, E1, DD, DL, F(c.wd0_pad), twin16(c.wd0_pad));
  ADDK_CHECK_LAUNCH();
  // regularisers (values for the log; their gradients are folded into the slab reduction): they depend on the weights
  // only, so they run here, next to the other chains, and not behind the join
  disc_weight_sumsq_kernel<<<dim3(64, 3), 256, 0, sd>>>(W + c.o_d_wl, E2, W + c.o_d_w0, (long long)E1 * DD, W + c.o_d_w1,
                                                        (long long)E2 * E1, stats);
  ADDK_CHECK_LAUNCH();
  TRY(gemm(sd, pr, F(c.dn), DL, 0, Wd0, DL, 1, e1, E1, R, E1, DL, W + c.o_d_b0, 1, nullptr, 0, 1, nullptr, nullptr, 0, F16_DROP_C));
  TRY(gemm(sd, pr, e1, E1, 0, W + c.o_d_w1, E1, 1, e2, E2, R, E2, E1, W + c.o_d_b1, 1));
  TRY(head1_forward(sd, e2, E2, R, E2, W + c.o_d_wl, W + c.o_d_bl, pred_d));
  disc_loss_kernel<<<(R + 255) / 256, 256, 0, sd>>>(pred_d, M, (float)c.disc_loss_weight, dpred_d, stats);
  ADDK_CHECK_LAUNCH();
  {
    size_t tot = (size_t)R * E2;
    uint16_t* const u2_16 = twin16(c.u2);
    uint16_t* const dh2_16 = twin16(dh2);
    const size_t nb = (tot + 255) / 256;
    disc_head_backward_kernel<<<(unsigned)(nb < 148 * 16 ? nb : 148 * 16), 256, 0, sd>>>(e2, W + c.o_d_wl, dpred_d, R, E2, F(c.u2),
                                                                            dh2, u2_16, dh2_16, amax_hook(sd, c.u2, R, E2, E2),
                                                                            amax_hook(sd, dh2, R, E2, E2));
    ADDK_CHECK_LAUNCH();
  }
  // input-gradient chain: u1 = m1 * (u2 W2), gx = u1 W1
  TRY(gemm(sd, pr, F(c.u2), E2, 0, W + c.o_d_w1, E1, 0, F(c.u1), E1, R, E1, E2, nullptr, 0, e1, E1, 1, nullptr, nullptr, 0, F16_DROP_C | F16_MASK));
  TRY(gemm(sd, pr, F(c.u1), E1, 0, Wd0, DL, 0, F(c.gx), DL, R, DL, E1));
  uint16_t* const dg16 = twin16(c.dg);
  grad_penalty_kernel<<<(R + 7) / 8, 256, 0, sd>>>(F(c.gx), M, R, DD, DL,
                                                   (float)(c.disc_loss_weight * c.disc_grad_penalty), F(c.dg), stats, dg16,
                                                   amax_hook(sd, c.dg, R, DL, DL));
  ADDK_CHECK_LAUNCH();
  // backward of the chain (second set of slabs)
  TRY(gemm(sd, pr, F(c.u1), E1, 1, F(c.dg), DL, 0, F(c.slabs) + (size_t)S * P + c.o_d_w0, DD, E1, DD, R, nullptr, 0,
           nullptr, 0, S, nullptr, nullptr, P));
  TRY(gemm(sd, pr, F(c.dg), DL, 0, Wd0, DL, 1, dv1, E1, R, E1, DL, nullptr, 0, e1, E1, 1, nullptr, nullptr, 0, F16_DROP_C | F16_MASK));
  TRY(gemm(sd, pr, F(c.u2), E2, 1, dv1, E1, 0, F(c.slabs) + (size_t)S * P + c.o_d_w1, E1, E2, E1, R, nullptr, 0, nullptr, 0,
           S, nullptr, nullptr, P));
  TRY(gemm(sd, pr, dv1, E1, 0, W + c.o_d_w1, E1, 1, du2, E2, R, E2, E1, nullptr, 0, e2, E2, 1, nullptr, nullptr, 0, 0, wd.colpart));
  if (wd.colpart && g_colpart_epi_for == du2 && g_colpart_epi_rows == (R + 31) / 32) {      // column sums left by the epilogue
    float* cwork = wd.colsum_work;
    unsigned int* ctick = (unsigned int*)(cwork + (size_t)COLSUM_CHUNKS * COLSUM_MAX_N) + 32;
    colsum_parts_kernel<<<dim3((E2 + 31) / 32, CP_CHUNKS), 256, 0, sd>>>(wd.colpart, g_colpart_epi_rows, E2,
                                                                          F(c.slabs) + (size_t)S * P + c.o_d_wl, P, S, cwork, ctick);
    ADDK_CHECK_LAUNCH();
    g_colpart_epi_for = nullptr;
  } else {
    TRY(colsum(sd, c, wd, du2, E2, R, E2, F(c.slabs) + (size_t)S * P + c.o_d_wl, nullptr));
  }
  // ordinary backward of the BCE terms
  TRY(wgrad(sd, c, wd, dpred_d, 1, e2, E2, R, 1, E2, c.o_d_wl, c.o_d_bl, 0));
  TRY(wgrad(sd, c, wd, dh2, E2, e1, E1, R, E2, E1, c.o_d_w1, c.o_d_b1, 0));
  TRY(gemm(sd, pr, dh2, E2, 0, W + c.o_d_w1, E1, 0, F(c.u1), E1, R, E1, E2, nullptr, 0, e1, E1, 1, nullptr, nullptr, 0, F16_DROP_C | F16_MASK, wd.colpart));
  TRY(wgrad(sd, c, wd, F(c.u1), E1, F(c.dn), DL, R, E1, DD, c.o_d_w0, c.o_d_b0, 0, true));
  if (multi) {
    cudaEventRecord(aux->join[0], sc);
    cudaEventRecord(aux->join[1], sd);
    cudaStreamWaitEvent(st, aux->join[0], 0);
    cudaStreamWaitEvent(st, aux->join[1], 0);
  }

  // ---------------- reduce slabs -> grads, AdamW, diagnostics ----------------
  SegTable t;
  t.P = P; t.n = 0;
  const long long offs[22] = {c.o_a_w0, c.o_a_b0, c.o_a_w1, c.o_a_b1, c.o_a_w2, c.o_a_b2, c.o_a_wm, c.o_a_bm,
                              c.o_c_w0, c.o_c_b0, c.o_c_w1, c.o_c_b1, c.o_c_w2, c.o_c_b2, c.o_c_wo, c.o_c_bo,
                              c.o_d_w0, c.o_d_b0, c.o_d_w1, c.o_d_b1, c.o_d_wl, c.o_d_bl};
  const float dlw = (float)c.disc_loss_weight;
  for (int k = 0; k < 22; ++k) {
    Seg& s = t.s[t.n++];
    s.begin = offs[k];
    s.end = (k + 1 < 22) ? offs[k + 1] : P;
    s.nslabs = S; s.l2 = 0.f;
    if (offs[k] == c.o_d_w0 || offs[k] == c.o_d_w1) { s.nslabs = 2 * S; s.l2 = dlw * 2.0f * (float)c.disc_weight_decay; }
    if (offs[k] == c.o_d_wl) { s.nslabs = 2 * S; s.l2 = dlw * 2.0f * (float)(c.disc_weight_decay + c.disc_logit_reg); }
  }
  const InfoArgs ia{stats, cnt, M, (float)c.action_bound_weight, (float)c.critic_loss_weight, dlw, (float)c.disc_logit_reg,
                    (float)c.disc_grad_penalty, (float)c.disc_weight_decay, F(c.info) + (size_t)step_index * 16};
  if (do_optim && !(c.grad_clip > 0.0) && addk_switches().fused_tail) {
    reduce_slabs_adamw_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(
        t, F(c.slabs), F(c.params), F(c.grads), F(c.exp_avg), F(c.exp_avg_sq),
        adamk_host(do_optim, c.lr, c.beta1, c.beta2, c.adam_eps, c.weight_decay, c.grad_scale), ia);
    ADDK_CHECK_LAUNCH();
    return ADDK_OK;
  }
  reduce_slabs_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(t, F(c.slabs), W, F(c.grads), ia);
  ADDK_CHECK_LAUNCH();
  if (do_optim) {
    if (c.grad_clip > 0.0)      // the coefficient lands in info[14] of this step (0 = clipping off)
      TRY(addk_clip_grad_norm(stream, F(c.grads), P, c.grad_clip, c.grad_scale, stats + 30, F(c.info) + (size_t)step_index * 16 + 14));
    TRY(addk_adamw(stream, F(c.params), F(c.grads), F(c.exp_avg), F(c.exp_avg_sq), P, do_optim, c.lr, c.beta1, c.beta2,
                   c.adam_eps, c.weight_decay, c.grad_scale));
  }
  return ADDK_OK;
}

// The derived copies of the parameters the inference entry points read: the 16-bit twin of the flat vector (fp16 hi / lo
// planes + max|x| word, or the bf16 copy) and the first-layer weights with padded rows.  Once per rollout / evaluation;
// contexts built with params16_current = 1 then skip these launches on every call.
extern "C" int addk_params_refresh(void* stream, void* ctx_host) {
  if (!ctx_host) return ADDK_ERR_ARG;
  const Ctx& c = *(const Ctx*)ctx_host;
  cudaStream_t st = (cudaStream_t)stream;
  const float* W = F(c.params);
  g_twin_ctx = &c;
  if (c.precision == 3 && c.params16) {
    f32_to_bf16_flat_kernel<<<148 * 4, 256, 0, st>>>(F(c.params), (uint16_t*)c.params16, c.num_params);
    ADDK_CHECK_LAUNCH();
  }
  if (c.precision == 4) {
    if (!c.params16 || !c.amax_slots) { addk_set_error("f16x3: the context has no parameter twin / max|x| slots"); return ADDK_ERR_ARG; }
    TRY(addk_f16x3_convert(st, W, 1, (int)c.num_params, (int)c.num_params, c.params16, c.num_params, (uint32_t*)c.amax_slots));
  }
  if (c.precision != 0) {
    const int OD = (int)c.obs_dim, OL = (int)c.obs_ld, DD = (int)c.disc_dim, DL = (int)c.disc_ld, H1 = (int)c.hid_a1, E1 = (int)c.hid_d1;
    if (OL > OD && c.wa0_pad) { pad_rows_kernel<<<(H1 * OL + 255) / 256, 256, 0, st>>>(W + c.o_a_w0, H1, OD, OL, F(c.wa0_pad), twin16(c.wa0_pad)); ADDK_CHECK_LAUNCH(); }
    if (OL > OD && c.wc0_pad) { pad_rows_kernel<<<(H1 * OL + 255) / 256, 256, 0, st>>>(W + c.o_c_w0, H1, OD, OL, F(c.wc0_pad), twin16(c.wc0_pad)); ADDK_CHECK_LAUNCH(); }
    if (c.wd0_pad) { pad_rows_kernel<<<(E1 * DL + 255) / 256, 256, 0, st>>>(W + c.o_d_w0, E1, DD, DL, F(c.wd0_pad), twin16(c.wd0_pad)); ADDK_CHECK_LAUNCH(); }
  }
  return ADDK_OK;
}

extern "C" int addk_actor_step(void* stream, void* ctx_host, const float* obs, const float* noise,
                               const float* exp_mask, int n, float* action, float* a_logp, float* obs_rec,
                               float* action_rec, float* logp_rec, float* mask_rec) {
  if (!ctx_host || !obs || !noise || !action || !a_logp || n <= 0) return ADDK_ERR_ARG;
  const Ctx& c = *(const Ctx*)ctx_host;
  if (n > c.mb_rows + 1) return ADDK_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int OD = (int)c.obs_dim, AD = (int)c.act_dim, AL = (int)c.act_ld, H3 = (int)c.hid_a3;
  const float* W = F(c.params);
  g_twin_ctx = &c;
  if (c.precision == 3 && c.params16 && !c.params16_current) {
    f32_to_bf16_flat_kernel<<<148 * 4, 256, 0, st>>>(F(c.params), (uint16_t*)c.params16, c.num_params);
    ADDK_CHECK_LAUNCH();
  }
  TRY(h3_params(c, st, 1));
  if (obs_rec) cudaMemcpyAsync(obs_rec, obs, (size_t)n * OD * sizeof(float), cudaMemcpyDeviceToDevice, st);
  TRY(trunk_forward(st, c, main_ws(c), obs, OD, OD, n, c.o_a_w0, c.o_a_b0, c.o_a_w1, c.o_a_b1, c.o_a_w2, c.o_a_b2, F(c.obs_mean),
                    F(c.obs_std), F(c.wa0_pad)));
  TRY(gemm(st, (int)c.precision, F(c.h3), H3, 0, W + c.o_a_wm, H3, 1, F(c.mean), AL, n, AD, H3, W + c.o_a_bm, 0));
  sample_action_kernel<<<(n + 7) / 8, 256, 0, st>>>(F(c.mean), AL, F(c.logstd), noise, exp_mask, F(c.a_mean), F(c.a_std),
                                                    n, AD, action, a_logp, action_rec, logp_rec, mask_rec);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}

extern "C" int addk_critic_eval(void* stream, void* ctx_host, const float* obs, long long n, float* vals) {
  if (!ctx_host || !obs || !vals || n <= 0) return ADDK_ERR_ARG;
  const Ctx& c = *(const Ctx*)ctx_host;
  cudaStream_t st = (cudaStream_t)stream;
  const int OD = (int)c.obs_dim, H3 = (int)c.hid_a3;
  const long long chunk = c.mb_rows + 1;
  const float* W = F(c.params);
  g_twin_ctx = &c;
  if (c.precision == 3 && c.params16 && !c.params16_current) {
    f32_to_bf16_flat_kernel<<<148 * 4, 256, 0, st>>>(F(c.params), (uint16_t*)c.params16, c.num_params);
    ADDK_CHECK_LAUNCH();
  }
  TRY(h3_params(c, st, 2));
  for (long long r0 = 0; r0 < n; r0 += chunk) {
    int rows = (int)((n - r0 < chunk) ? n - r0 : chunk);
    TRY(trunk_forward(st, c, main_ws(c), obs + r0 * OD, OD, OD, rows, c.o_c_w0, c.o_c_b0, c.o_c_w1, c.o_c_b1, c.o_c_w2, c.o_c_b2,
                      F(c.obs_mean), F(c.obs_std), F(c.wc0_pad)));
    TRY(head1_forward(st, F(c.h3), H3, rows, H3, W + c.o_c_wo, W + c.o_c_bo, vals + r0));
  }
  return ADDK_OK;
}

extern "C" int addk_disc_eval(void* stream, void* ctx_host, const float* disc_obs, const float* disc_obs_demo,
                              long long n, float* logits) {
  if (!ctx_host || !disc_obs || !disc_obs_demo || !logits || n <= 0) return ADDK_ERR_ARG;
  const Ctx& c = *(const Ctx*)ctx_host;
  cudaStream_t st = (cudaStream_t)stream;
  const int DD = (int)c.disc_dim, DL = (int)c.disc_ld, E1 = (int)c.hid_d1, E2 = (int)c.hid_d2, pr = (int)c.precision;
  const long long chunk = c.mb_rows;
  const float* W = F(c.params);
  float *e1 = F(c.h1), *e2 = F(c.h3);
  g_twin_ctx = &c;
  if (c.precision == 3 && c.params16 && !c.params16_current) {
    f32_to_bf16_flat_kernel<<<148 * 4, 256, 0, st>>>(F(c.params), (uint16_t*)c.params16, c.num_params);
    ADDK_CHECK_LAUNCH();
  }
  TRY(h3_params(c, st, 3));
  if (!c.params16_current) {
    pad_rows_kernel<<<(E1 * DL + 255) / 256, 256, 0, st>>>(W + c.o_d_w0, E1, DD, DL, F(c.wd0_pad), twin16(c.wd0_pad));
    ADDK_CHECK_LAUNCH();
  }
  for (long long r0 = 0; r0 < n; r0 += chunk) {
    int rows = (int)((n - r0 < chunk) ? n - r0 : chunk);
    long long tot = (long long)rows * DL;
    diff_normalize_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(disc_obs + r0 * DD, disc_obs_demo + r0 * DD,
                                                                        F(c.disc_mean_abs), rows, DD, DL, F(c.gx), twin16(c.gx));
    ADDK_CHECK_LAUNCH();
    TRY(gemm(st, pr, F(c.gx), DL, 0, F(c.wd0_pad), DL, 1, e1, E1, rows, E1, DL, W + c.o_d_b0, 1, nullptr, 0, 1, nullptr, nullptr, 0, F16_DROP_C));
    TRY(gemm(st, pr, e1, E1, 0, W + c.o_d_w1, E1, 1, e2, E2, rows, E2, E1, W + c.o_d_b1, 1));
    TRY(head1_forward(st, e2, E2, rows, E2, W + c.o_d_wl, W + c.o_d_bl, logits + r0));
  }
  return ADDK_OK;
}
