// Reference-motion table: build (one-time, per clip) and runtime gather.
//
// Replaces the reference's MotionLib._extract_frame_data / _load_motion_pkl velocity block /
// _precompute_motion_steps / calc_motion_frame / get_precomputed_motion_step
// (add_gym/anim/motion_lib.py:100-131,164-335) and KinCharModel.dof_to_rot / rot_to_dof /
// compute_frame_dof_vel (add_gym/anim/kin_char_model.py:196-266,595-639).
//
// HBM layout of the 100 Hz step table: one 288-byte row per step,
//     [ root_pos 3 | root_rot 4 (wxyz) | dof_pos D ]  [ root_vel 3 | root_ang_vel 3 | dof_vel D | pad ]
//       `---------------- pose half ---------------'    `---------------- velocity half -------------'
// with D = 29 -> 36 + 36 = 72 floats, so both halves start 16-byte aligned and a consumer that needs
// only the pose (target / discriminator-demo rows) reads 9 float4 instead of six scattered gathers.
#include "common.cuh"
#include "addk.h"

namespace addk {

// ------------------------------------------------------------------------------------------------
// Stage A: per source frame (30 fps) quantities.
//   frames   [F, 7+D] fp32, file layout: pos xyz, quat xyzw, D hinge angles in file column order
//   jrot     [F, D, 4]  quat_pos(axis_angle_to_quat(axis_d, dof))      (motion_lib.py:113-114)
//   fvel     [F, 6+D]   root_vel(3) root_ang_vel(3) dof_vel(D)          (motion_lib.py:203-215)
// One thread per (frame, dof); dof index D is the "root" worker.
// ------------------------------------------------------------------------------------------------
__global__ void frame_prep_kernel(const float* __restrict__ frames, int F, int D,
                                  const int* __restrict__ col_of_dof, const float* __restrict__ dof_axis,
                                  float fps, float inv_fps_as_dt, float* __restrict__ jrot,
                                  float* __restrict__ fvel) {
  int f = blockIdx.x;
  int d = threadIdx.x;
  if (f >= F || d > D) return;
  const int W = 7 + D;
  // the last frame repeats the velocity of frame F-2 (motion_lib.py:205,212; kin_char_model.py:231-233)
  int f0 = (f < F - 1) ? f : (F - 2);
  if (f0 < 0) f0 = 0;
  int f1 = (F > 1) ? f0 + 1 : f0;
  if (d < D) {
    Vec3 ax = {dof_axis[3 * d], dof_axis[3 * d + 1], dof_axis[3 * d + 2]};
    int col = 7 + col_of_dof[d];
    Quat q = quat_pos(axis_angle_to_quat(ax, frames[(size_t)f * W + col]));
    float* o = jrot + ((size_t)f * D + d) * 4;
    o[0] = q.w; o[1] = q.x; o[2] = q.y; o[3] = q.z;
    Quat q0 = quat_pos(axis_angle_to_quat(ax, frames[(size_t)f0 * W + col]));
    Quat q1 = quat_pos(axis_angle_to_quat(ax, frames[(size_t)f1 * W + col]));
    // drot = normalize(conj(q0) * q1); vel = exp_map(drot) / dt projected on the hinge axis
    Quat dr = quat_normalize(quat_mul(quat_conj(q0), q1));
    Vec3 e = quat_to_exp_map(dr);
    e.x = e.x / inv_fps_as_dt; e.y = e.y / inv_fps_as_dt; e.z = e.z / inv_fps_as_dt;
    float v = add_rn(add_rn(mul_rn(ax.x, e.x), mul_rn(ax.y, e.y)), mul_rn(ax.z, e.z));
    fvel[(size_t)f * (6 + D) + 6 + d] = (F > 1) ? v : 0.0f;
  } else {
    const float* a = frames + (size_t)f0 * W;
    const float* b = frames + (size_t)f1 * W;
    float* o = fvel + (size_t)f * (6 + D);
    if (F > 1) {
      o[0] = mul_rn(fps, sub_rn(b[0], a[0]));
      o[1] = mul_rn(fps, sub_rn(b[1], a[1]));
      o[2] = mul_rn(fps, sub_rn(b[2], a[2]));
      Quat r0 = {a[6], a[3], a[4], a[5]}, r1 = {b[6], b[3], b[4], b[5]};  // xyzw -> wxyz (motion_lib.py:10-15)
      Vec3 e = quat_to_exp_map(quat_diff(r0, r1));
      o[3] = mul_rn(fps, e.x); o[4] = mul_rn(fps, e.y); o[5] = mul_rn(fps, e.z);
    } else {
      for (int i = 0; i < 6; ++i) o[i] = 0.0f;
    }
  }
}

// torch.arange(0, len, dt) on the CPU (fp32 result): ATen fills blocks of 2x8 lanes as
//   float(double(start) + step*idx_block) + lane*step   (lane term in double, AVX2 Vectorized<float>),
// and the last n % 16 elements with the scalar expression float(start + step*i).  The reference builds
// its step times with exactly that call (motion_lib.py:296-298), so the table is sampled at these times.
__device__ __forceinline__ float arange_time(int i, int n, double dt) {
  int tail_begin = n - (n % 16);
  if (i >= tail_begin) return (float)(dt * (double)i);
  int blk = i - (i % 8);
  float base = (float)(dt * (double)blk);
  return (float)((double)base + (double)(i % 8) * dt);
}

// ------------------------------------------------------------------------------------------------
// Stage B: one row of the 100 Hz table per step (motion_lib.py:61-88,118-150,361-372).
// One warp-sized group per step: lane d < D blends joint d, lane D blends the root.
// ------------------------------------------------------------------------------------------------
__global__ void step_table_kernel(const float* __restrict__ frames, const float* __restrict__ jrot,
                                  const float* __restrict__ fvel, int F, int D, const float* __restrict__ dof_axis,
                                  int n_steps, double dt, float motion_len, int loop_wrap,
                                  float* __restrict__ table, int row_stride, long long row0,
                                  float* __restrict__ joint_rot_out, long long* __restrict__ frame_idx_out) {
  int s = blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
  int d = threadIdx.x & 31;
  if (s >= n_steps || d > D) return;
  const int W = 7 + D;
  const int half = (7 + D + 3) & ~3;  // pose half, padded to a float4 boundary (36 for D=29)
  float t = arange_time(s, n_steps, dt);
  // calc_phase (motion_lib.py:361-372)
  float phase = t / motion_len;
  float wraps = 0.0f;
  if (loop_wrap) { wraps = floorf(phase); phase = sub_rn(phase, wraps); }
  phase = fminf(fmaxf(phase, 0.0f), 1.0f);
  // _calc_frame_blend (motion_lib.py:118-131)
  float nf1 = (float)(F - 1);
  float pf = mul_rn(phase, nf1);
  long long i0 = (long long)pf;
  long long i1 = (i0 + 1 < F - 1) ? i0 + 1 : (long long)(F - 1);
  float blend = sub_rn(pf, (float)i0);
  float* row = table + (size_t)(row0 + s) * row_stride;
  if (d < D) {
    const float* a = jrot + ((size_t)i0 * D + d) * 4;
    const float* b = jrot + ((size_t)i1 * D + d) * 4;
    Quat q = slerp({a[0], a[1], a[2], a[3]}, {b[0], b[1], b[2], b[3]}, blend);
    Vec3 ax = {dof_axis[3 * d], dof_axis[3 * d + 1], dof_axis[3 * d + 2]};
    row[7 + d] = quat_twist_angle(q, ax);                 // KinCharModel.rot_to_dof, hinge
    row[half + 6 + d] = fvel[(size_t)i0 * (6 + D) + 6 + d];  // velocities are taken un-blended from frame i0
    if (joint_rot_out) {
      float* o = joint_rot_out + ((size_t)(row0 + s) * D + d) * 4;
      o[0] = q.w; o[1] = q.x; o[2] = q.y; o[3] = q.z;
    }
  } else {
    const float* a = frames + (size_t)i0 * W;
    const float* b = frames + (size_t)i1 * W;
    float om = sub_rn(1.0f, blend);
    float px = add_rn(mul_rn(om, a[0]), mul_rn(blend, b[0]));
    float py = add_rn(mul_rn(om, a[1]), mul_rn(blend, b[1]));
    float pz = add_rn(mul_rn(om, a[2]), mul_rn(blend, b[2]));
    if (loop_wrap) {  // _calc_loop_offset (motion_lib.py:133-150): floor(t/len) * (pos[-1]-pos[0]) with z zeroed
      const float* l = frames + (size_t)(F - 1) * W;
      px = add_rn(px, mul_rn(wraps, sub_rn(l[0], frames[0])));
      py = add_rn(py, mul_rn(wraps, sub_rn(l[1], frames[1])));
    }
    Quat r = slerp({a[6], a[3], a[4], a[5]}, {b[6], b[3], b[4], b[5]}, blend);
    row[0] = px; row[1] = py; row[2] = pz;
    row[3] = r.w; row[4] = r.x; row[5] = r.y; row[6] = r.z;
    const float* v = fvel + (size_t)i0 * (6 + D);
    for (int i = 0; i < 6; ++i) row[half + i] = v[i];
    for (int i = 7 + D; i < half; ++i) row[i] = 0.0f;
    for (int i = half + 6 + D; i < row_stride; ++i) row[i] = 0.0f;
    if (frame_idx_out) { frame_idx_out[2 * (row0 + s)] = i0; frame_idx_out[2 * (row0 + s) + 1] = i1; }
  }
}

// ------------------------------------------------------------------------------------------------
// MotionLib.calc_motion_frame (motion_lib.py:61-88) at ARBITRARY (clip, time) queries: the run-time form of the
// interpolation the table build applies at the 100 Hz grid -- root lerp, root / joint slerp, twist angle, velocities
// of frame i0, loop offset of WRAP clips.  The 30 fps source frames of all clips sit concatenated in HBM
// (frames / jrot / fvel as produced by stage A); `frame_start` is the cumulative 30 fps frame count, which is
// what the reference's `_motion_start_idx` really indexes (motion_lib.py:118-131).  One warp per query.
// ------------------------------------------------------------------------------------------------
__global__ void motion_frame_kernel(const float* __restrict__ frames, const float* __restrict__ jrot,
                                    const float* __restrict__ fvel, const long long* __restrict__ frame_start,
                                    const long long* __restrict__ num_frames, const float* __restrict__ lengths,
                                    const int* __restrict__ loop_modes, int D, const float* __restrict__ dof_axis,
                                    const long long* __restrict__ ids, const float* __restrict__ times, int n,
                                    float* __restrict__ root_pos, float* __restrict__ root_rot, float* __restrict__ root_vel,
                                    float* __restrict__ root_ang, float* __restrict__ joint_rot, float* __restrict__ dof_pos,
                                    float* __restrict__ dof_vel) {
  const int i = blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
  const int d = threadIdx.x & 31;
  if (i >= n || d > D) return;
  const int W = 7 + D;
  const long long m = ids[i];
  const long long F = num_frames[m], f0 = frame_start[m];
  const float t = times[i], len = lengths[m];
  const bool wrap = loop_modes[m] == 1;
  float phase = t / len;                                       // calc_phase (motion_lib.py:361-372)
  if (wrap) phase = sub_rn(phase, floorf(phase));
  phase = fminf(fmaxf(phase, 0.0f), 1.0f);
  const float nf1 = (float)(F - 1);                            // _calc_frame_blend (motion_lib.py:118-131)
  const float pf = mul_rn(phase, nf1);
  long long i0 = (long long)pf;
  long long i1 = (i0 + 1 < F - 1) ? i0 + 1 : F - 1;
  const float blend = sub_rn(pf, (float)i0);
  i0 += f0; i1 += f0;
  if (d < D) {
    const float* a = jrot + ((size_t)i0 * D + d) * 4;
    const float* b = jrot + ((size_t)i1 * D + d) * 4;
    const Quat q = slerp({a[0], a[1], a[2], a[3]}, {b[0], b[1], b[2], b[3]}, blend);
    if (joint_rot) { float* o = joint_rot + ((size_t)i * D + d) * 4; o[0] = q.w; o[1] = q.x; o[2] = q.y; o[3] = q.z; }
    if (dof_pos) {
      const Vec3 ax = {dof_axis[3 * d], dof_axis[3 * d + 1], dof_axis[3 * d + 2]};
      dof_pos[(size_t)i * D + d] = quat_twist_angle(q, ax);
    }
    if (dof_vel) dof_vel[(size_t)i * D + d] = fvel[(size_t)i0 * (6 + D) + 6 + d];
  } else {
    const float* a = frames + (size_t)i0 * W;
    const float* b = frames + (size_t)i1 * W;
    const float om = sub_rn(1.0f, blend);
    float px = add_rn(mul_rn(om, a[0]), mul_rn(blend, b[0]));
    float py = add_rn(mul_rn(om, a[1]), mul_rn(blend, b[1]));
    const float pz = add_rn(mul_rn(om, a[2]), mul_rn(blend, b[2]));
    if (wrap) {                                                // _calc_loop_offset (motion_lib.py:133-150)
      const float wraps = floorf(t / len);
      const float* first = frames + (size_t)f0 * W;
      const float* last = frames + (size_t)(f0 + F - 1) * W;
      px = add_rn(px, mul_rn(wraps, sub_rn(last[0], first[0])));
      py = add_rn(py, mul_rn(wraps, sub_rn(last[1], first[1])));
    }
    if (root_pos) { root_pos[(size_t)i * 3] = px; root_pos[(size_t)i * 3 + 1] = py; root_pos[(size_t)i * 3 + 2] = pz; }
    if (root_rot) {
      const Quat r = slerp({a[6], a[3], a[4], a[5]}, {b[6], b[3], b[4], b[5]}, blend);
      float* o = root_rot + (size_t)i * 4; o[0] = r.w; o[1] = r.x; o[2] = r.y; o[3] = r.z;
    }
    const float* v = fvel + (size_t)i0 * (6 + D);
    if (root_vel) for (int k = 0; k < 3; ++k) root_vel[(size_t)i * 3 + k] = v[k];
    if (root_ang) for (int k = 0; k < 3; ++k) root_ang[(size_t)i * 3 + k] = v[3 + k];
  }
}

// ------------------------------------------------------------------------------------------------
// Runtime lookup (motion_lib.py:322-335):
//   frame = trunc(time * dt_inv); frame = clip(frame, 0, S_total-1); idx = frame + start_idx[motion_id]
// `start_idx` is whatever the caller's MotionLib holds -- by default the reference's cumulative sum of
// 30 fps source frame counts (quirk Q2), kept so that indices are bit-identical.  The row actually
// read is additionally clamped into the table for memory safety (the reference would raise).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long motion_row(float time, float dt_inv, long long s_total, long long start) {
  long long fr = (long long)mul_rn(time, dt_inv);
  fr = fr < 0 ? 0 : (fr > s_total - 1 ? s_total - 1 : fr);
  return fr + start;
}

__global__ void motion_gather_kernel(const float* __restrict__ table, int row_stride, int D, long long s_total,
                                     const long long* __restrict__ start_idx, float dt_inv,
                                     const long long* __restrict__ ids, const float* __restrict__ times, int n,
                                     float* __restrict__ root_pos, float* __restrict__ root_rot,
                                     float* __restrict__ root_vel, float* __restrict__ root_ang,
                                     float* __restrict__ dof_pos, float* __restrict__ dof_vel,
                                     long long* __restrict__ idx_out) {
  int i = blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
  int lane = threadIdx.x & 31;
  if (i >= n) return;
  long long idx = motion_row(times[i], dt_inv, s_total, start_idx[ids[i]]);
  if (idx_out && lane == 0) idx_out[i] = idx;
  long long r = idx < 0 ? 0 : (idx > s_total - 1 ? s_total - 1 : idx);
  const float* row = table + (size_t)r * row_stride;
  const int half = (7 + D + 3) & ~3;
  for (int c = lane; c < 7 + D; c += 32) {
    float v = row[c];
    if (c < 3) { if (root_pos) root_pos[(size_t)i * 3 + c] = v; }
    else if (c < 7) { if (root_rot) root_rot[(size_t)i * 4 + (c - 3)] = v; }
    else if (dof_pos) dof_pos[(size_t)i * D + (c - 7)] = v;
  }
  for (int c = lane; c < 6 + D; c += 32) {
    float v = row[half + c];
    if (c < 3) { if (root_vel) root_vel[(size_t)i * 3 + c] = v; }
    else if (c < 6) { if (root_ang) root_ang[(size_t)i * 3 + (c - 3)] = v; }
    else if (dof_vel) dof_vel[(size_t)i * D + (c - 6)] = v;
  }
}

}  // namespace addk

using namespace addk;

extern "C" int addk_motion_table_build(void* stream, const float* frames, int num_frames, int num_dofs,
                                       const int* col_of_dof, const float* dof_axis, float fps, float frame_dt,
                                       int n_steps, double dt, float motion_len, int loop_wrap,
                                       float* jrot_work, float* fvel_work, float* table, int row_stride,
                                       long long row0, float* joint_rot_out, long long* frame_idx_out) {
  if (!frames || !table || !jrot_work || !fvel_work || num_frames < 1 || num_dofs < 1 || num_dofs > 31)
    return ADDK_ERR_ARG;
  const int half = (7 + num_dofs + 3) & ~3;
  if (row_stride < 2 * half) return ADDK_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  frame_prep_kernel<<<num_frames, 32, 0, st>>>(frames, num_frames, num_dofs, col_of_dof, dof_axis, fps, frame_dt,
                                                jrot_work, fvel_work);
  ADDK_CHECK_LAUNCH();
  if (n_steps > 0) {
    int wpb = 8;
    step_table_kernel<<<(n_steps + wpb - 1) / wpb, wpb * 32, 0, st>>>(
        frames, jrot_work, fvel_work, num_frames, num_dofs, dof_axis, n_steps, dt, motion_len, loop_wrap, table,
        row_stride, row0, joint_rot_out, frame_idx_out);
    ADDK_CHECK_LAUNCH();
  }
  return ADDK_OK;
}

extern "C" int addk_motion_frame(void* stream, const float* frames, const float* jrot, const float* fvel,
                                 const long long* frame_start, const long long* num_frames, const float* lengths,
                                 const int* loop_modes, int num_dofs, const float* dof_axis, const long long* ids,
                                 const float* times, int n, float* root_pos, float* root_rot, float* root_vel,
                                 float* root_ang_vel, float* joint_rot, float* dof_pos, float* dof_vel) {
  if (n == 0) return ADDK_OK;
  if (!frames || !jrot || !fvel || !frame_start || !num_frames || !lengths || !loop_modes || !dof_axis || !ids || !times ||
      n < 0 || num_dofs < 1 || num_dofs > 31)
    return ADDK_ERR_ARG;
  const int wpb = 8;
  motion_frame_kernel<<<(n + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(
      frames, jrot, fvel, frame_start, num_frames, lengths, loop_modes, num_dofs, dof_axis, ids, times, n, root_pos,
      root_rot, root_vel, root_ang_vel, joint_rot, dof_pos, dof_vel);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}

extern "C" int addk_motion_gather(void* stream, const float* table, int row_stride, int num_dofs,
                                  long long s_total, const long long* start_idx, float dt_inv,
                                  const long long* ids, const float* times, int n, float* root_pos,
                                  float* root_rot, float* root_vel, float* root_ang_vel, float* dof_pos,
                                  float* dof_vel, long long* idx_out) {
  if (n == 0) return ADDK_OK;
  if (!table || !start_idx || !ids || !times || n < 0) return ADDK_ERR_ARG;
  int wpb = 8;
  motion_gather_kernel<<<(n + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(
      table, row_stride, num_dofs, s_total, start_idx, dt_inv, ids, times, n, root_pos, root_rot, root_vel,
      root_ang_vel, dof_pos, dof_vel, idx_out);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}
