// Shared device helpers for the add-gym B200 hot path: wxyz quaternion math that follows the
// reference's formulas operation by operation (add_gym/util/torch_util.py:34-406), warp/block
// reductions, 128-bit load/store helpers and the status codes of the C-ABI.
//
// All quaternions are (w,x,y,z).  Functions that feed integer results (frame indices, done flags)
// use the explicit round-to-nearest intrinsics so that nvcc cannot contract a*b+c into an FMA --
// the reference computes every elementwise op with its own fp32 rounding.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <atomic>

#define ADDK_OK 0
#define ADDK_ERR_ARG 1
#define ADDK_ERR_LAUNCH 2
#define ADDK_ERR_UNSUPPORTED 3

// Call right after every kernel launch: counts it (addk_launch_count) and surfaces launch errors.
#define ADDK_CHECK_LAUNCH()                                                                  \
  do {                                                                                       \
    g_addk_launches.fetch_add(1, std::memory_order_relaxed);                                 \
    cudaError_t e__ = cudaGetLastError();                                                    \
    if (e__ != cudaSuccess) { addk_set_error(cudaGetErrorString(e__)); return ADDK_ERR_LAUNCH; } \
  } while (0)

void addk_set_error(const char* msg);
extern std::atomic<long long> g_addk_launches;

namespace addk {

struct Quat { float w, x, y, z; };
struct Vec3 { float x, y, z; };

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }

// a*b - c*d style expressions are written with separately rounded products on purpose.
__device__ __forceinline__ Vec3 cross3(const Vec3& a, const Vec3& b) {
  Vec3 r;
  r.x = sub_rn(mul_rn(a.y, b.z), mul_rn(a.z, b.y));
  r.y = sub_rn(mul_rn(a.z, b.x), mul_rn(a.x, b.z));
  r.z = sub_rn(mul_rn(a.x, b.y), mul_rn(a.y, b.x));
  return r;
}

// torch_util.py:65-71  v + q_w * t + cross(q_v, t),  t = 2 * cross(q_v, v)
__device__ __forceinline__ Vec3 quat_rotate(const Quat& q, const Vec3& v) {
  Vec3 qv = {q.x, q.y, q.z};
  Vec3 t = cross3(qv, v);
  t.x = mul_rn(2.0f, t.x); t.y = mul_rn(2.0f, t.y); t.z = mul_rn(2.0f, t.z);
  Vec3 c = cross3(qv, t);
  Vec3 r;
  r.x = add_rn(add_rn(v.x, mul_rn(q.w, t.x)), c.x);
  r.y = add_rn(add_rn(v.y, mul_rn(q.w, t.y)), c.y);
  r.z = add_rn(add_rn(v.z, mul_rn(q.w, t.z)), c.z);
  return r;
}

// torch_util.py:230-242  [rotate(q, e_x), rotate(q, e_z)]
__device__ __forceinline__ void quat_to_tan_norm(const Quat& q, float* out6) {
  Vec3 ex = {1.0f, 0.0f, 0.0f}, ez = {0.0f, 0.0f, 1.0f};
  Vec3 t = quat_rotate(q, ex);
  Vec3 n = quat_rotate(q, ez);
  out6[0] = t.x; out6[1] = t.y; out6[2] = t.z; out6[3] = n.x; out6[4] = n.y; out6[5] = n.z;
}

// torch_util.py:48-62
__device__ __forceinline__ Quat quat_mul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = sub_rn(sub_rn(sub_rn(mul_rn(a.w, b.w), mul_rn(a.x, b.x)), mul_rn(a.y, b.y)), mul_rn(a.z, b.z));
  r.x = sub_rn(add_rn(add_rn(mul_rn(a.w, b.x), mul_rn(a.x, b.w)), mul_rn(a.y, b.z)), mul_rn(a.z, b.y));
  r.y = add_rn(add_rn(sub_rn(mul_rn(a.w, b.y), mul_rn(a.x, b.z)), mul_rn(a.y, b.w)), mul_rn(a.z, b.x));
  r.z = add_rn(sub_rn(add_rn(mul_rn(a.w, b.z), mul_rn(a.x, b.y)), mul_rn(a.y, b.x)), mul_rn(a.z, b.w));
  return r;
}

__device__ __forceinline__ Quat quat_conj(const Quat& q) { return {q.w, -q.x, -q.y, -q.z}; }

// torch_util.py:40-45  flip the sign when w < 0
__device__ __forceinline__ Quat quat_pos(const Quat& q) {
  float s = (q.w < 0.0f) ? -1.0f : 1.0f;
  return {s * q.w, s * q.x, s * q.y, s * q.z};
}

__device__ __forceinline__ float norm3(float x, float y, float z) {
  return sqrtf(add_rn(add_rn(mul_rn(x, x), mul_rn(y, y)), mul_rn(z, z)));
}
__device__ __forceinline__ float norm4(const Quat& q) {
  return sqrtf(add_rn(add_rn(add_rn(mul_rn(q.w, q.w), mul_rn(q.x, q.x)), mul_rn(q.y, q.y)), mul_rn(q.z, q.z)));
}

// torch_util.py:11-14 / 34-37   x / clamp(norm(x), 1e-9)
__device__ __forceinline__ Quat quat_unit(const Quat& q) {
  float n = fmaxf(norm4(q), 1e-9f);
  return {q.w / n, q.x / n, q.y / n, q.z / n};
}
__device__ __forceinline__ Quat quat_normalize(const Quat& q) { return quat_unit(quat_pos(q)); }

// torch_util.py:74-94
__device__ __forceinline__ void quat_to_axis_angle(const Quat& q_in, Vec3& axis, float& angle) {
  Quat q = quat_pos(q_in);
  float len = norm3(q.x, q.y, q.z);
  float a = mul_rn(2.0f, atan2f(len, q.w));
  bool ok = len > 1e-5f;
  angle = ok ? a : 0.0f;
  axis.x = ok ? q.x / len : 0.0f;
  axis.y = ok ? q.y / len : 0.0f;
  axis.z = ok ? q.z / len : 1.0f;
}

// torch_util.py:176-182   quat_unit([cos(a/2), normalize(axis) * sin(a/2)])
__device__ __forceinline__ Quat axis_angle_to_quat(const Vec3& axis, float angle) {
  float th = angle / 2.0f;
  float n = fmaxf(norm3(axis.x, axis.y, axis.z), 1e-9f);
  float s = sinf(th), c = cosf(th);
  Quat q = {c, mul_rn(axis.x / n, s), mul_rn(axis.y / n, s), mul_rn(axis.z / n, s)};
  return quat_unit(q);
}

// torch_util.py:203-208
__device__ __forceinline__ Vec3 quat_to_exp_map(const Quat& q) {
  Vec3 ax; float ang;
  quat_to_axis_angle(q, ax, ang);
  return {mul_rn(ang, ax.x), mul_rn(ang, ax.y), mul_rn(ang, ax.z)};
}

// torch_util.py:269-284   dq = q1 * conj(q0)
__device__ __forceinline__ Quat quat_diff(const Quat& q0, const Quat& q1) { return quat_mul(q1, quat_conj(q0)); }
__device__ __forceinline__ float quat_diff_angle(const Quat& q0, const Quat& q1) {
  Vec3 ax; float ang;
  quat_to_axis_angle(quat_diff(q0, q1), ax, ang);
  return ang;
}

// torch_util.py:300-323
__device__ __forceinline__ Quat slerp(const Quat& q0, Quat q1, float t) {
  float c = add_rn(add_rn(add_rn(mul_rn(q0.w, q1.w), mul_rn(q0.x, q1.x)), mul_rn(q0.y, q1.y)), mul_rn(q0.z, q1.z));
  if (c < 0.0f) { q1.w = -q1.w; q1.x = -q1.x; q1.y = -q1.y; q1.z = -q1.z; }
  c = fabsf(c);
  float half = acosf(c);
  float s = sqrtf(sub_rn(1.0f, mul_rn(c, c)));
  float ra = sinf(mul_rn(sub_rn(1.0f, t), half)) / s;
  float rb = sinf(mul_rn(t, half)) / s;
  Quat r = {add_rn(mul_rn(ra, q0.w), mul_rn(rb, q1.w)), add_rn(mul_rn(ra, q0.x), mul_rn(rb, q1.x)),
            add_rn(mul_rn(ra, q0.y), mul_rn(rb, q1.y)), add_rn(mul_rn(ra, q0.z), mul_rn(rb, q1.z))};
  if (fabsf(s) < 0.001f) {
    r.w = add_rn(mul_rn(0.5f, q0.w), mul_rn(0.5f, q1.w));
    r.x = add_rn(mul_rn(0.5f, q0.x), mul_rn(0.5f, q1.x));
    r.y = add_rn(mul_rn(0.5f, q0.y), mul_rn(0.5f, q1.y));
    r.z = add_rn(mul_rn(0.5f, q0.z), mul_rn(0.5f, q1.z));
  }
  if (fabsf(c) >= 1.0f) r = q0;
  return r;
}

// torch_util.py:385-406   signed rotation angle of q about `axis`
__device__ __forceinline__ float quat_twist_angle(const Quat& q, const Vec3& axis) {
  float p = add_rn(add_rn(mul_rn(axis.x, q.x), mul_rn(axis.y, q.y)), mul_rn(axis.z, q.z));
  Quat tw = {q.w, mul_rn(p, axis.x), mul_rn(p, axis.y), mul_rn(p, axis.z)};
  tw = quat_normalize(tw);
  Vec3 ax; float ang;
  quat_to_axis_angle(tw, ax, ang);
  float d = add_rn(add_rn(mul_rn(axis.x, ax.x), mul_rn(axis.y, ax.y)), mul_rn(axis.z, ax.z));
  return (d < 0.0f) ? -ang : ang;
}

// torch_util.py:326-356   rotation about +z that undoes the heading of q
__device__ __forceinline__ Quat calc_heading_quat_inv(const Quat& q) {
  Vec3 ex = {1.0f, 0.0f, 0.0f};
  Vec3 d = quat_rotate(q, ex);
  float heading = atan2f(d.y, d.x);
  Vec3 ez = {0.0f, 0.0f, 1.0f};
  return axis_angle_to_quat(ez, -heading);
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of one double per thread; result valid in thread 0.  blockDim.x <= 1024.
__device__ __forceinline__ double block_sum(double v, double* smem32) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem32[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    r = (lane < nw) ? smem32[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

// fp32 -> bf16, round to nearest even
__device__ __forceinline__ uint16_t to_bf16(float x) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(0.f), "f"(x));
  return (uint16_t)(r & 0xFFFFu);
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stg4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
// streaming (evict-first) 128-bit store for write-once experience rows
__device__ __forceinline__ void stg4_cs(float* p, const float4& v) { __stcs(reinterpret_cast<float4*>(p), v); }

}  // namespace addk
