// Returns, advantages, discriminator reward, running statistics and AdamW: the HBM-bound
// elementwise / reduction stages of one training iteration.
//
// Replaces (reference add_gym/learning/):
//   compute_td_lambda_return + next_vals masking + adv   base_agent.py:624-647, ppo_agent.py:126-146
//   advantage std_mean / normalise / clamp                 ppo_agent.py:147-153
//   AMPAgent._calc_disc_rewards tail, reward mix           amp_agent.py:201-205, add_agent.py:124-133
//   Normalizer.record/update, DiffNormalizer.record/update normalizer.py:25-80, diff_normalizer.py:24-45
//   torch.optim.AdamW (single tensor semantics)            mp_optimizer.py:38
// Reductions accumulate in fp64 and are combined with fp64 atomics (two numbers per block), so they are
// at least as accurate as the reference's fp32 cascade sums; elementwise formulas keep its op order.
#include "common.cuh"
#include "addk.h"
#include "adam.cuh"

namespace addk {

// thread per env, serial over T (reverse), coalesced over N.  The recurrence only chains `next_ret`: the four loads of
// eight time steps are issued before the first dependent add, so a thread keeps 32 loads in flight instead of 4.
__global__ void __launch_bounds__(64) td_lambda_kernel(const float* __restrict__ reward, const float* __restrict__ next_vals,
                                 const float* __restrict__ vals, const int32_t* __restrict__ done, int T, int N,
                                 float discount, float td_lambda, float succ_val, float fail_val,
                                 float* __restrict__ tar_val, float* __restrict__ adv) {
  constexpr int U = 8;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float next_ret = 0.f;
  for (int t0 = T - 1; t0 >= 0; t0 -= U) {
    float nv[U], r[U], vl[U];
    int d[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t0 - u;
      if (t >= 0) {
        const size_t i = (size_t)t * N + n;
        nv[u] = __ldg(next_vals + i); d[u] = __ldg(done + i); r[u] = __ldg(reward + i); vl[u] = __ldg(vals + i);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t0 - u;
      if (t < 0) break;
      const size_t i = (size_t)t * N + n;
      float nvu = nv[u];
      if (d[u] == 2) nvu = succ_val;
      if (d[u] == 1) nvu = fail_val;
      float ret;
      if (t == T - 1) {
        ret = add_rn(r[u], mul_rn(discount, nvu));
      } else {
        const float reset = (d[u] != 0) ? 1.0f : 0.0f;
        const float lam = mul_rn(td_lambda, sub_rn(1.0f, reset));
        ret = add_rn(r[u], mul_rn(discount, add_rn(mul_rn(sub_rn(1.0f, lam), nvu), mul_rn(lam, next_ret))));
      }
      tar_val[i] = ret;
      adv[i] = sub_rn(ret, vl[u]);
      next_ret = ret;
    }
  }
}

__global__ void masked_moments_kernel(const float* __restrict__ x, const float* __restrict__ mask, int n,
                                      double* __restrict__ work3) {
  __shared__ double sm[32];
  double s = 0.0, s2 = 0.0, c = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (!mask || mask[i] == 1.0f) { double v = x[i]; s += v; s2 += v * v; c += 1.0; }
  }
  s = block_sum(s, sm); s2 = block_sum(s2, sm); c = block_sum(c, sm);
  if (threadIdx.x == 0) { atomicAdd(work3, s); atomicAdd(work3 + 1, s2); atomicAdd(work3 + 2, c); }
}

__device__ __forceinline__ void mean_std_unbiased(const double* w, float& mean, float& sd) {
  double c = w[2], m = (c > 0) ? w[0] / c : 0.0;
  double var = (c > 1) ? (w[1] - c * m * m) / (c - 1.0) : 0.0;
  mean = (float)m;
  sd = (float)sqrt(var > 0.0 ? var : 0.0);
}

__global__ void adv_normalize_kernel(float* __restrict__ adv, int n, float clip, const double* __restrict__ work3,
                                     float* __restrict__ stats_out) {
  float mean, sd;
  mean_std_unbiased(work3, mean, sd);
  float den = fmaxf(sd, 1e-5f);
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && stats_out) { stats_out[0] = mean; stats_out[1] = sd; }
  if (i >= n) return;
  float v = sub_rn(adv[i], mean) / den;
  adv[i] = fminf(fmaxf(v, -clip), clip);
}

__global__ void disc_reward_kernel(const float* __restrict__ logits, float* __restrict__ reward, int n, float scale,
                                   float w_task, float w_disc, double* __restrict__ work3) {
  __shared__ double sm[32];
  double s = 0.0, s2 = 0.0, c = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float prob = 1.0f / add_rn(1.0f, expf(-logits[i]));
    float dr = mul_rn(-logf(fmaxf(sub_rn(1.0f, prob), 0.0001f)), scale);
    reward[i] = add_rn(mul_rn(w_task, reward[i]), mul_rn(w_disc, dr));
    s += dr; s2 += (double)dr * dr; c += 1.0;
  }
  s = block_sum(s, sm); s2 = block_sum(s2, sm); c = block_sum(c, sm);
  if (threadIdx.x == 0) { atomicAdd(work3, s); atomicAdd(work3 + 1, s2); atomicAdd(work3 + 2, c); }
}

__global__ void moments_finalize_kernel(const double* __restrict__ work3, float* __restrict__ stats_out) {
  float mean, sd;
  mean_std_unbiased(work3, mean, sd);
  stats_out[0] = mean; stats_out[1] = sd;
}

// Column sums over a [n, dim] row-major matrix; block = 256 threads over columns, ROWS rows per block.
template <int MODE>
__global__ void column_stats_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, int dim,
                                    int rows_per_block, double* __restrict__ out) {
  long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block < n ? r0 + rows_per_block : n;
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
    double s = 0.0, s2 = 0.0;
    for (long long r = r0; r < r1; ++r) {
      float v = a[r * dim + c];
      if (MODE == 0) { s += v; s2 += (double)v * v; }
      else { s += fabsf(sub_rn(v, b[r * dim + c])); }
    }
    atomicAdd(out + c, s);
    if (MODE == 0) atomicAdd(out + dim + c, s2);
  }
}

__global__ void normalizer_update_kernel(const double* __restrict__ sums, double new_count, int dim,
                                         int64_t* count, float* mean, float* mean_sq, float* std_, float min_var) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  long long old = count[0];
  long long total = old + (long long)new_count;
  if (c < dim) {
    float nm = (float)(sums[c] / new_count), nms = (float)(sums[dim + c] / new_count);
    float w_old = (float)old / (float)total, w_new = (float)new_count / (float)total;
    float m = add_rn(mul_rn(w_old, mean[c]), mul_rn(w_new, nm));
    float ms = add_rn(mul_rn(w_old, mean_sq[c]), mul_rn(w_new, nms));
    mean[c] = m; mean_sq[c] = ms;
    std_[c] = sqrtf(fmaxf(sub_rn(ms, mul_rn(m, m)), min_var));
  }
  __syncthreads();
  // every block needs the old count; only after all blocks read it may it change -> single block launch
  if (c == 0) count[0] = total;
}

__global__ void diff_normalizer_update_kernel(const double* __restrict__ sum_abs, double new_count, int dim,
                                              int64_t* count, float* mean_abs) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  long long old = count[0];
  long long total = old + (long long)new_count;
  if (c < dim) {
    float nm = (float)(sum_abs[c] / new_count);
    float w_old = (float)old / (float)total, w_new = (float)new_count / (float)total;
    mean_abs[c] = add_rn(mul_rn(w_old, mean_abs[c]), mul_rn(w_new, nm));
  }
  __syncthreads();
  if (c == 0) count[0] = total;
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, const AdamK k) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float pi = p[i], mi = m[i], vi = v[i];
  adam1(k, pi, g[i], mi, vi);
  p[i] = pi; m[i] = mi; v[i] = vi;
}
// 16-byte aligned vectors: four parameters per thread (7 x 128-bit accesses), the n % 4 tail by the thread after the last quad
__global__ void __launch_bounds__(256) adamw_vec4_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, long long n, const AdamK k) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n4 = n >> 2;
  if (i < n4) {
    float4 P = *reinterpret_cast<const float4*>(p + 4 * i), M = *reinterpret_cast<const float4*>(m + 4 * i),
           V = *reinterpret_cast<const float4*>(v + 4 * i);
    const float4 G = ldg4(g + 4 * i);
    adam1(k, P.x, G.x, M.x, V.x); adam1(k, P.y, G.y, M.y, V.y); adam1(k, P.z, G.z, M.z, V.z); adam1(k, P.w, G.w, M.w, V.w);
    stg4(p + 4 * i, P); stg4(m + 4 * i, M); stg4(v + 4 * i, V);
  } else if (i == n4) {
    for (long long j = 4 * n4; j < n; ++j) {
      float pj = p[j], mj = m[j], vj = v[j];
      adam1(k, pj, g[j], mj, vj);
      p[j] = pj; m[j] = mj; v[j] = vj;
    }
  }
}

}  // namespace addk

using namespace addk;

extern "C" int addk_td_lambda(void* stream, const float* reward, const float* next_vals, const float* vals,
                              const int32_t* done, int T, int N, float discount, float td_lambda, float succ_val,
                              float fail_val, float* tar_val, float* adv) {
  if (!reward || !next_vals || !vals || !done || !tar_val || !adv || T <= 0 || N <= 0) return ADDK_ERR_ARG;
  td_lambda_kernel<<<(N + 63) / 64, 64, 0, (cudaStream_t)stream>>>(reward, next_vals, vals, done, T, N, discount,
                                                                     td_lambda, succ_val, fail_val, tar_val, adv);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}

// ExperienceBuffer._sample_rand_idx (experience_buffer.py:90-113): window of the permutation, modulo the sample count
__global__ void perm_window_kernel(const long long* __restrict__ perm, long long len, long long head, int n, long long count,
                                   long long* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long j = head + i;
  if (j >= len) j -= len;                     // (n <= len: one wrap at most)
  long long v = perm[j] % count;              // torch.remainder: the sign of the divisor (count > 0, perm >= 0)
  out[i] = v < 0 ? v + count : v;
}

extern "C" int addk_perm_window(void* stream, const long long* perm, long long perm_len, long long head, int n,
                                long long sample_count, long long* out_idx) {
  if (!perm || !out_idx || n <= 0 || perm_len <= 0 || n > perm_len || head < 0 || head > perm_len || sample_count <= 0)
    return ADDK_ERR_ARG;
  perm_window_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(perm, perm_len, head, n, sample_count, out_idx);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}

extern "C" int addk_adv_normalize(void* stream, float* adv, const float* rand_action_mask, int n, float clip,
                                  double* work3, float* stats_out) {
  if (!adv || !work3 || n <= 0) return ADDK_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(work3, 0, 3 * sizeof(double), st);
  int bl = (n + 255) / 256; if (bl > 1184) bl = 1184;
  masked_moments_kernel<<<bl, 256, 0, st>>>(adv, rand_action_mask, n, work3);
  ADDK_CHECK_LAUNCH();
  adv_normalize_kernel<<<(n + 255) / 256, 256, 0, st>>>(adv, n, clip, work3, stats_out);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}

extern "C" int addk_disc_reward(void* stream, const float* logits, float* reward_inout, int n, float scale,
                                float w_task, float w_disc, double* work3, float* stats_out) {
  if (!logits || !reward_inout || !work3 || n <= 0) return ADDK_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(work3, 0, 3 * sizeof(double), st);
  int bl = (n + 255) / 256; if (bl > 1184) bl = 1184;
  disc_reward_kernel<<<bl, 256, 0, st>>>(logits, reward_inout, n, scale, w_task, w_disc, work3);
  ADDK_CHECK_LAUNCH();
  if (stats_out) { moments_finalize_kernel<<<1, 1, 0, st>>>(work3, stats_out); ADDK_CHECK_LAUNCH(); }
  return ADDK_OK;
}

extern "C" int addk_column_stats(void* stream, const float* a, const float* b, long long n, int dim, int mode,
                                 double* out) {
  if (!a || !out || n <= 0 || dim <= 0 || (mode == 1 && !b)) return ADDK_ERR_ARG;
  const int rows = 128;
  int bl = (int)((n + rows - 1) / rows);
  if (mode == 0) column_stats_kernel<0><<<bl, 256, 0, (cudaStream_t)stream>>>(a, b, n, dim, rows, out);
  else column_stats_kernel<1><<<bl, 256, 0, (cudaStream_t)stream>>>(a, b, n, dim, rows, out);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}

extern "C" int addk_normalizer_update(void* stream, const double* sums, double new_count, int dim, int64_t* count,
                                      float* mean, float* mean_sq, float* std_, float min_var) {
  if (!sums || !count || !mean || !mean_sq || !std_ || dim <= 0 || dim > 1024) return ADDK_ERR_ARG;
  if (new_count <= 0) return ADDK_OK;  // normalizer.py:62-63
  normalizer_update_kernel<<<1, ((dim + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(sums, new_count, dim, count, mean,
                                                                                    mean_sq, std_, min_var);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}

extern "C" int addk_diff_normalizer_update(void* stream, const double* sum_abs, double new_count, int dim,
                                           int64_t* count, float* mean_abs) {
  if (!sum_abs || !count || !mean_abs || dim <= 0 || dim > 1024) return ADDK_ERR_ARG;
  diff_normalizer_update_kernel<<<1, ((dim + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(sum_abs, new_count, dim,
                                                                                         count, mean_abs);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}

extern "C" int addk_adamw(void* stream, float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                          long long n, int step, double lr, double beta1, double beta2, double eps,
                          double weight_decay, double grad_scale) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || n <= 0 || step < 1) return ADDK_ERR_ARG;
  const AdamK k = adamk_host(step, lr, beta1, beta2, eps, weight_decay, grad_scale);
  const bool aligned = ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
                         reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0;
  if (aligned)
    adamw_vec4_kernel<<<(unsigned)(((n >> 2) + 1 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, k);
  else
    adamw_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, k);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}
