// Fused per-env work around the physics step: reference-motion window gather, history push,
// policy observation, agent / demo discriminator observations, tracking reward, done flags,
// episode bookkeeping and the experience-buffer row -- one kernel, one warp per environment.
//
// Replaces (reference add_gym/learning/add/):
//   ADDObservation.update_motion / compute_obs      add_observation.py:163-306,356-419,422-717
//   ADDReward.compute_reward                          add_reward.py:54-177
//   ADDDone.compute_done + Manipulator contact bool   add_done.py:59-147, robot.py:221-231
//   ADDAgent._record_data_post_step                   add_agent.py:93-108
//   ReturnTracker.update                              base_agent.py:596-621
//   ADDObservation.reset_idx / _reset_disc_hist       add_observation.py:308-344
//   ADDMotion.sample_time, AdaptiveSegmentSampler     add_motion.py:53-61, sampler.py:20-92
//
// Data movement per env-step (D=29, defaults): the warp stages its 9-row window of the step table
// (288-byte rows, float4 loads), the simulator state and the two older history rows in shared memory,
// assembles the 264/114/114-float output rows there, and streams them out with 128-bit stores.
// HBM-bound: ~2.0 KB read + ~3.6 KB written per env-step.
#include "common.cuh"
#include "addk.h"
#include "switches.h"

namespace addk {

constexpr int WPB = 4;  // warps (= envs) per block

enum { F_ADVANCE = 1, F_UPDATE_MOTION = 2, F_REWARD_DONE = 4, F_MASKED = 8 };

// Manipulator.get_ground_contact_forces_v2 (robot.py:221-231): does any VALID contact of env e involve a link that may
// not touch the plane?  The contact list is [N, contact_slots] with the engine's CURRENT width (MJWarp returns width 0
// while nothing touches, Genesis pads to the per-step maximum): lanes stride over the slots from `first`; the caller
// combines the lanes with __any_sync.
__device__ __forceinline__ int contact_hit(const addk_sim_state& sim, uint64_t noncontact_mask, int e, int lane, int first) {
  int hit = 0;
  if (sim.contact_link_masks) {          // per-env link bitmasks (addk_contact_link_mask): isin(link_a | link_b, noncontact ids)
    if (first == 0 && lane == 0)
      hit = ((sim.contact_link_masks[2 * (size_t)e] | sim.contact_link_masks[2 * (size_t)e + 1]) & noncontact_mask) != 0ull;
    return hit;
  }
  if (sim.valid) {
    for (int c = first + lane; c < sim.contact_slots; c += 32) {
      const size_t ci = (size_t)e * sim.ld_contact + c;
      if (sim.valid[ci]) {
        const int la = sim.link_a[ci], lb = sim.link_b[ci];
        const bool ha = la >= 0 && la < 64 && ((noncontact_mask >> la) & 1ull);
        const bool hb = lb >= 0 && lb < 64 && ((noncontact_mask >> lb) & 1ull);
        hit |= (ha || hb) ? 1 : 0;
      }
    }
  }
  return hit;
}

struct StepParams {
  addk_task task;
  addk_motion_lib lib;
  addk_sim_state sim;
  addk_env_buffers env;
  addk_exp_row exp;
  const float* dof_err_w;
  const uint8_t* env_mask;
  int n, newest_slot, flags, has_exp;
};

__device__ __forceinline__ long long table_row(const addk_motion_lib& lib, float time, float dt_inv, long long start) {
  long long fr = (long long)mul_rn(time, dt_inv);
  fr = fr < 0 ? 0 : (fr > lib.s_total - 1 ? lib.s_total - 1 : fr);
  long long idx = fr + start;
  return idx < 0 ? 0 : (idx > lib.s_total - 1 ? lib.s_total - 1 : idx);
}

__device__ __forceinline__ Quat ldq(const float* p) { return {p[0], p[1], p[2], p[3]}; }

// pose-obs of one history / demo step into dst: [pos3 | tan_norm 6 | dof D | (vel3 ang3 dofvel D)]
// src is a staged row: pose half at src[0..], velocity half at src[half..]   (add_observation.py:462-554)
__device__ __forceinline__ void disc_step_small(const float* src, int half, float* dst, int D, bool global_obs,
                                                bool vel_obs) {
  dst[0] = global_obs ? src[0] : 0.0f;
  dst[1] = global_obs ? src[1] : 0.0f;
  dst[2] = src[2];
  Quat q = ldq(src + 3);
  quat_to_tan_norm(q, dst + 3);
  if (vel_obs) {
    float* v = dst + 9 + D;
    Vec3 lv = {src[half], src[half + 1], src[half + 2]}, av = {src[half + 3], src[half + 4], src[half + 5]};
    if (!global_obs) {
      Quat h = calc_heading_quat_inv(q);
      lv = quat_rotate(h, lv);
      av = quat_rotate(h, av);
    }
    v[0] = lv.x; v[1] = lv.y; v[2] = lv.z; v[3] = av.x; v[4] = av.y; v[5] = av.z;
  }
}

// FAST = true is the instance for the reference's default task (configs/task/pose.yaml: 29 dofs, 6 target steps,
// 3 discriminator steps, global obs with root height, no velocity / phase obs): every loop bound and row size is a
// compile-time constant, which removes ~half of the kernel's instructions (index arithmetic; the kernel is
// issue-bound, not HBM-bound, see profiles/).  FAST = false reads the same quantities from the task struct.
template <bool FAST>
__global__ void __launch_bounds__(WPB * 32) env_step_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(16) float smem[];
  const addk_task& tk = p.task;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * WPB + warp;
  if (e >= p.n) return;
  if ((p.flags & F_MASKED) && !p.env_mask[e]) return;

  const int D = FAST ? 29 : tk.num_dofs;
  const int half = (7 + D + 3) & ~3;
  const int RS = FAST ? 72 : p.lib.row_stride;
  const int nT = FAST ? 6 : (tk.enable_tar_obs ? tk.num_tar_steps : 0);
  const int nH = FAST ? 3 : tk.num_disc_steps;
  const bool vel = FAST ? false : (tk.enable_vel_obs != 0), glob = FAST ? true : (tk.global_obs != 0);
  const bool root_h = FAST ? true : (tk.root_height_obs != 0);
  const bool phase_obs = FAST ? false : (tk.enable_phase_obs != 0);
  const int obs_dim = FAST ? 264 : tk.obs_dim, disc_dim = FAST ? 114 : tk.disc_obs_dim;
  const int obs_pad = (obs_dim + 3) & ~3, disc_pad = (disc_dim + 3) & ~3;
  // per-warp shared layout
  const int per_warp = RS /*ref*/ + nT * half + nH * RS /*demo*/ + RS /*sim*/ + nH * RS /*hist*/ + obs_pad + 2 * disc_pad;
  float* s_ref = smem + (size_t)warp * per_warp;
  float* s_tar = s_ref + RS;
  float* s_demo = s_tar + nT * half;
  float* s_sim = s_demo + nH * RS;
  float* s_hist = s_sim + RS;
  float* s_obs = s_hist + nH * RS;
  float* s_disc = s_obs + obs_pad;
  float* s_dobs = s_disc + disc_pad;

  // ---- time and table rows ---------------------------------------------------------------------
  // lane 0 owns the read-modify-write of time_buf; everyone else gets the value by shuffle
  float t = 0.f;
  if (lane == 0) {
    t = p.env.time_buf[e];
    if (p.flags & F_ADVANCE) {
      t = add_rn(t, tk.ctrl_dt);
      p.env.time_buf[e] = t;
    }
  }
  t = __shfl_sync(0xffffffffu, t, 0);
  const long long mid = p.env.motion_ids[e];
  const float mt = add_rn(t, p.env.motion_time_offsets[e]);
  const long long start = p.lib.start_idx[mid];

  {  // reference row at mt: full row
    const float* src = p.lib.table + (size_t)table_row(p.lib, mt, tk.dt_inv, start) * RS;
    for (int i = lane; i < RS / 4; i += 32) stg4(s_ref + 4 * i, ldg4(src + 4 * i));
  }
  {  // target rows at mt + dt*k : pose half only
    const int h4 = half / 4;
    for (int i = lane; i < nT * h4; i += 32) {
      int k = i / h4, c = i - k * h4;
      const float* src = p.lib.table + (size_t)table_row(p.lib, add_rn(mt, tk.tar_offsets[k]), tk.dt_inv, start) * RS;
      stg4(s_tar + k * half + 4 * c, ldg4(src + 4 * c));
    }
  }
  {  // demo rows at mt + disc_offsets[j] (oldest -> newest)
    const int w4 = (vel ? RS : half) / 4;
    for (int i = lane; i < nH * w4; i += 32) {
      int j = i / w4, c = i - j * w4;
      const float* src = p.lib.table + (size_t)table_row(p.lib, add_rn(mt, tk.disc_offsets[j]), tk.dt_inv, start) * RS;
      stg4(s_demo + j * RS + 4 * c, ldg4(src + 4 * c));
    }
  }
  // ---- simulator state -> packed row ---------------------------------------------------------
  if (lane < D) {
    s_sim[7 + lane] = p.sim.dof_pos[(size_t)e * p.sim.ld_dof_pos + lane];
    s_sim[half + 6 + lane] = p.sim.dof_vel[(size_t)e * p.sim.ld_dof_vel + lane];
  }
  if (lane < 3) {
    s_sim[lane] = p.sim.root_pos[(size_t)e * p.sim.ld_root_pos + lane];
    s_sim[half + lane] = p.sim.root_vel[(size_t)e * p.sim.ld_root_vel + lane];
    s_sim[half + 3 + lane] = p.sim.root_ang[(size_t)e * p.sim.ld_root_ang + lane];
  }
  if (lane < 4) s_sim[3 + lane] = p.sim.root_rot[(size_t)e * p.sim.ld_root_rot + lane];
  if (lane == 31) {
    for (int i = 7 + D; i < half; ++i) s_sim[i] = 0.0f;
    for (int i = half + 6 + D; i < RS; ++i) s_sim[i] = 0.0f;
  }
  // ---- history ring: logical j (oldest..newest) lives in slot (newest_slot + 1 + j) % nH -------------
  float* g_hist = p.env.hist + (size_t)e * nH * p.env.hist_stride;
  const bool push = (p.flags & F_UPDATE_MOTION) != 0;
  {
    const int r4 = RS / 4;
    for (int i = lane; i < nH * r4; i += 32) {
      int j = i / r4, c = i - j * r4;
      if (push && j == nH - 1) continue;  // newest entry is the state being pushed
      int slot = (p.newest_slot + 1 + j) % nH;
      stg4(s_hist + j * RS + 4 * c, *reinterpret_cast<const float4*>(g_hist + (size_t)slot * p.env.hist_stride + 4 * c));
    }
  }
  __syncwarp();
  if (push) {
    float* dst = g_hist + (size_t)p.newest_slot * p.env.hist_stride;
    for (int i = lane; i < RS / 4; i += 32) {
      float4 v = *reinterpret_cast<const float4*>(s_sim + 4 * i);
      stg4(dst + 4 * i, v);
      stg4(s_hist + (nH - 1) * RS + 4 * i, v);
    }
    // _update_ref_motion: ref_* <- table row (add_observation.py:163-175)
    if (lane < 3) {
      p.env.ref_root_pos[(size_t)e * 3 + lane] = s_ref[lane];
      p.env.ref_root_vel[(size_t)e * 3 + lane] = s_ref[half + lane];
      p.env.ref_root_ang_vel[(size_t)e * 3 + lane] = s_ref[half + 3 + lane];
    }
    if (lane < 4) p.env.ref_root_rot[(size_t)e * 4 + lane] = s_ref[3 + lane];
    if (lane < D) {
      p.env.ref_dof_pos[(size_t)e * D + lane] = s_ref[7 + lane];
      p.env.ref_dof_vel[(size_t)e * D + lane] = s_ref[half + 6 + lane];
    }
  }
  __syncwarp();

  // ---- policy observation (add_observation.py:422-459,578-717) ----------------------------------
  const Quat root_q = ldq(s_sim + 3);
  int o = 0;
  const int o_h = o;            o += root_h ? 1 : 0;
  const int o_rot = o;          o += 6;
  const int o_dof = o;          o += D;
  const int o_vel = o;          o += vel ? 6 + D : 0;
  const int o_phase = o;        o += phase_obs ? 1 + 2 * tk.num_phase_encoding : 0;
  const int o_tar = o;
  const int tar_pos_dim = root_h ? 3 : 2;
  const int tar_dim = tar_pos_dim + 6 + D;
  if (lane == 31) {
    if (root_h) s_obs[o_h] = s_sim[2];
    Quat hinv = {1.f, 0.f, 0.f, 0.f};
    if (!glob) hinv = calc_heading_quat_inv(root_q);
    Quat q = glob ? root_q : quat_mul(hinv, root_q);
    quat_to_tan_norm(q, s_obs + o_rot);
    if (vel) {
      Vec3 lv = {s_sim[half], s_sim[half + 1], s_sim[half + 2]}, av = {s_sim[half + 3], s_sim[half + 4], s_sim[half + 5]};
      if (!glob) { lv = quat_rotate(hinv, lv); av = quat_rotate(hinv, av); }
      float* v = s_obs + o_vel;
      v[0] = lv.x; v[1] = lv.y; v[2] = lv.z; v[3] = av.x; v[4] = av.y; v[5] = av.z;
    }
    if (phase_obs) {  // calc_phase + compute_phase_obs (motion_lib.py:361-372; add_observation.py:557-575)
      float ph = mt / p.lib.lengths[mid];
      if (p.lib.loop_modes[mid] == 1) ph = sub_rn(ph, floorf(ph));
      ph = fminf(fmaxf(ph, 0.0f), 1.0f);
      s_obs[o_phase] = ph;
      for (int k = 0; k < tk.num_phase_encoding; ++k) {
        float sc = mul_rn(mul_rn(2.0f, 3.14159274101257324f), exp2f((float)k));
        float v = mul_rn(ph, sc);
        s_obs[o_phase + 1 + k] = sinf(v);
        s_obs[o_phase + 1 + tk.num_phase_encoding + k] = cosf(v);
      }
    }
  }
  if (lane < D) {
    s_obs[o_dof + lane] = s_sim[7 + lane];
    if (vel) s_obs[o_vel + 6 + lane] = s_sim[half + 6 + lane];
  }
  if (lane < nT) {  // compute_tar_obs, one target step per lane
    const float* tr = s_tar + lane * half;
    float* dst = s_obs + o_tar + lane * tar_dim;
    Vec3 ref_pos = glob ? Vec3{s_sim[0], s_sim[1], s_sim[2]} : Vec3{s_tar[0], s_tar[1], s_tar[2]};
    Vec3 dp = {sub_rn(tr[0], ref_pos.x), sub_rn(tr[1], ref_pos.y), sub_rn(tr[2], ref_pos.z)};
    Quat q = ldq(tr + 3);
    if (!glob) {
      Quat hinv = calc_heading_quat_inv(ldq(s_tar + 3));
      dp = quat_rotate(hinv, dp);
      q = quat_mul(hinv, q);
    }
    dst[0] = dp.x; dst[1] = dp.y;
    if (root_h) dst[2] = tr[2];
    quat_to_tan_norm(q, dst + tar_pos_dim);
  }
  for (int k = 0; k < nT; ++k)
    if (lane < D) s_obs[o_tar + k * tar_dim + tar_pos_dim + 6 + lane] = s_tar[k * half + 7 + lane];

  // ---- discriminator observations: agent history and demo rows (add_observation.py:276-294,356-419)
  const int step_dim = 9 + D + (vel ? 6 + D : 0);
  if (lane < nH) disc_step_small(s_hist + lane * RS, half, s_disc + lane * step_dim, D, glob, vel);
  else if (lane >= 16 && lane < 16 + nH)
    disc_step_small(s_demo + (lane - 16) * RS, half, s_dobs + (lane - 16) * step_dim, D, glob, vel);
  for (int j = 0; j < nH; ++j) {
    if (lane < D) {
      s_disc[j * step_dim + 9 + lane] = s_hist[j * RS + 7 + lane];
      s_dobs[j * step_dim + 9 + lane] = s_demo[j * RS + 7 + lane];
      if (vel) {
        s_disc[j * step_dim + 15 + D + lane] = s_hist[j * RS + half + 6 + lane];
        s_dobs[j * step_dim + 15 + D + lane] = s_demo[j * RS + half + 6 + lane];
      }
    }
  }
  __syncwarp();

  // ---- stream the rows out ----------------------------------------------------------------------
  {
    float* g = p.env.obs_buf + (size_t)e * obs_dim;
    float* gx = p.has_exp ? p.exp.next_obs + (size_t)e * obs_dim : nullptr;
    if ((obs_dim & 3) == 0) {
      for (int i = lane; i < obs_dim / 4; i += 32) {
        float4 v = *reinterpret_cast<const float4*>(s_obs + 4 * i);
        stg4(g + 4 * i, v);
        if (gx) stg4_cs(gx + 4 * i, v);
      }
    } else {
      for (int i = lane; i < obs_dim; i += 32) { g[i] = s_obs[i]; if (gx) gx[i] = s_obs[i]; }
    }
    float* gd = p.env.disc_obs + (size_t)e * disc_dim;
    float* gm = p.env.disc_obs_demo + (size_t)e * disc_dim;
    float* xd = p.has_exp ? p.exp.disc_obs + (size_t)e * disc_dim : nullptr;
    float* xm = p.has_exp ? p.exp.disc_obs_demo + (size_t)e * disc_dim : nullptr;
    if ((disc_dim & 1) == 0) {  // rows are 8-byte aligned: 64-bit stores
      for (int i = lane; i < disc_dim / 2; i += 32) {
        float2 a = *reinterpret_cast<const float2*>(s_disc + 2 * i);
        float2 b = *reinterpret_cast<const float2*>(s_dobs + 2 * i);
        *reinterpret_cast<float2*>(gd + 2 * i) = a;
        *reinterpret_cast<float2*>(gm + 2 * i) = b;
        if (xd) { __stcs(reinterpret_cast<float2*>(xd + 2 * i), a); __stcs(reinterpret_cast<float2*>(xm + 2 * i), b); }
      }
    } else {
      for (int i = lane; i < disc_dim; i += 32) {
        gd[i] = s_disc[i]; gm[i] = s_dobs[i];
        if (xd) { xd[i] = s_disc[i]; xm[i] = s_dobs[i]; }
      }
    }
  }
  if (p.has_exp && lane == 0) {
    p.exp.motion_ids[e] = mid;
    p.exp.motion_times[e] = mt;
  }
  if (!(p.flags & F_REWARD_DONE)) return;

  // ---- tracking reward (add_reward.py:104-177) --------------------------------------------------
  float pe = 0.f, ve = 0.f, de = 0.f;
  if (lane < D) {
    float w = p.dof_err_w[lane];
    float pd = sub_rn(s_ref[7 + lane], s_sim[7 + lane]);
    float vd = sub_rn(s_ref[half + 6 + lane], s_sim[half + 6 + lane]);
    pe = mul_rn(mul_rn(w, pd), pd);
    ve = mul_rn(mul_rn(w, vd), vd);
    de = mul_rn(pd, pd);
  }
  pe = warp_sum(pe); ve = warp_sum(ve); de = warp_sum(de);
  int contact = contact_hit(p.sim, tk.noncontact_link_mask, e, lane, 0);
  contact = __any_sync(0xffffffffu, contact);
  if (lane == 0) {
    Vec3 rp = {s_sim[0], s_sim[1], s_sim[2]}, tp = {s_ref[0], s_ref[1], s_ref[2]};
    Vec3 dpos = {sub_rn(tp.x, rp.x), sub_rn(tp.y, rp.y), sub_rn(tp.z, rp.z)};
    Vec3 dpos_r = dpos;
    if (!tk.track_root) { dpos_r.x = 0.f; dpos_r.y = 0.f; }
    if (!tk.track_root_h) dpos_r.z = 0.f;
    float root_pos_err = add_rn(add_rn(mul_rn(dpos_r.x, dpos_r.x), mul_rn(dpos_r.y, dpos_r.y)), mul_rn(dpos_r.z, dpos_r.z));
    Quat rq = root_q, tq = ldq(s_ref + 3);
    Vec3 rv = {s_sim[half], s_sim[half + 1], s_sim[half + 2]}, ra = {s_sim[half + 3], s_sim[half + 4], s_sim[half + 5]};
    Vec3 tv = {s_ref[half], s_ref[half + 1], s_ref[half + 2]}, ta = {s_ref[half + 3], s_ref[half + 4], s_ref[half + 5]};
    if (!tk.track_root) {  // convert_to_local_root (add_reward.py:91-102)
      Quat h0 = calc_heading_quat_inv(rq), h1 = calc_heading_quat_inv(tq);
      rv = quat_rotate(h0, rv); ra = quat_rotate(h0, ra); rq = quat_mul(h0, rq);
      tv = quat_rotate(h1, tv); ta = quat_rotate(h1, ta); tq = quat_mul(h1, tq);
    }
    float rot_err = quat_diff_angle(rq, tq);
    rot_err = mul_rn(rot_err, rot_err);
    Vec3 dv = {sub_rn(tv.x, rv.x), sub_rn(tv.y, rv.y), sub_rn(tv.z, rv.z)};
    Vec3 da = {sub_rn(ta.x, ra.x), sub_rn(ta.y, ra.y), sub_rn(ta.z, ra.z)};
    float vel_err = add_rn(add_rn(mul_rn(dv.x, dv.x), mul_rn(dv.y, dv.y)), mul_rn(dv.z, dv.z));
    float ang_err = add_rn(add_rn(mul_rn(da.x, da.x), mul_rn(da.y, da.y)), mul_rn(da.z, da.z));
    float pose_r = expf(mul_rn(-tk.pose_scale, pe));
    float vel_r = expf(mul_rn(-tk.vel_scale, ve));
    float root_pose_r = expf(mul_rn(-tk.root_pose_scale, add_rn(root_pos_err, mul_rn(0.1f, rot_err))));
    float root_vel_r = expf(mul_rn(-tk.root_vel_scale, add_rn(vel_err, mul_rn(0.1f, ang_err))));
    float r = add_rn(add_rn(add_rn(mul_rn(tk.pose_w, pose_r), mul_rn(tk.vel_w, vel_r)), mul_rn(tk.root_pose_w, root_pose_r)),
                     mul_rn(tk.root_vel_w, root_vel_r));
    // ---- done flags (add_done.py:97-147): later writes win -------------------------------------
    int done = 0;
    if (t >= tk.ep_len) done = 3;
    if (mt >= p.lib.lengths[mid] && p.lib.loop_modes[mid] != 1) done = 2;
    if (tk.enable_early_termination) {
      bool failed = contact != 0;
      if (tk.pose_termination) {
        bool pf = (de / (float)D) > tk.pose_termination_dist;
        if (tk.track_root) {
          float re = add_rn(add_rn(mul_rn(dpos.x, dpos.x), mul_rn(dpos.y, dpos.y)), mul_rn(dpos.z, dpos.z));
          pf = pf || (re > tk.pose_termination_dist);
        }
        failed = failed || pf;
      }
      if (failed && t > 0.0f) done = 1;
    }
    p.env.reward[e] = r;
    p.env.done[e] = done;
    if (p.has_exp) { p.exp.reward[e] = r; p.exp.done[e] = done; }
    // ---- ReturnTracker.update ---------------------------------------------------------------------
    if (p.env.return_buf) {
      float ret = add_rn(p.env.return_buf[e], r);
      long long len = p.env.ep_len_buf[e] + 1;
      if (done != 0) {
        atomicAdd(p.env.tracker_sums, (double)ret);
        atomicAdd(p.env.tracker_sums + 1, (double)len);
        atomicAdd(reinterpret_cast<unsigned long long*>(p.env.tracker_count), 1ull);
        p.env.eps_per_env[e] += 1;
        ret = 0.f; len = 0;
      }
      p.env.return_buf[e] = ret;
      p.env.ep_len_buf[e] = len;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Lean instance of the step kernel for the reference's default task (configs/task/pose.yaml: 29 dofs, 6 target steps,
// 3 discriminator steps, global obs with root height, no velocity / phase obs).  Same arithmetic, op for op, as
// env_step_kernel above; what changes is the bookkeeping: every offset is a compile-time constant, the 13 quaternion
// -> tangent/normal conversions run in 13 lanes at once, the row assembly is a handful of unrolled shared-memory copies.
// The generic kernel is issue-bound (2,220 warp-instructions per env); this one is sized to let HBM be the limit.
//
// Shared memory per warp (floats): rows[9][36] pose halves of the table rows (0 = ref at t, 1..6 = targets, 7..8 =
// demo t-0.02 / t-0.01), refv[36] velocity half of the ref row, sim[72], hist[3][36] (oldest..newest), obs[264],
// disc[116], demo[116].
// ------------------------------------------------------------------------------------------------
namespace fast {
constexpr int D = 29, HALF = 36, RS = 72, NT = 6, NH = 3, OBS = 264, DISC = 114, STEP = 38;
constexpr int O_ROWS = 0, O_REFV = 9 * HALF, O_SIM = O_REFV + HALF, O_HIST = O_SIM + RS, O_OBS = O_HIST + NH * HALF,
              O_DISC = O_OBS + OBS, O_DEMO = O_DISC + 116, PER_WARP = O_DEMO + 116;
}  // namespace fast

struct FastTail { float pe, ve, de, t, mt, m_len, ret0; int contact, m_loop; long long len0; };

// ---- bulk-copy staging (cp.async.bulk + mbarrier): one 144-byte copy per table / history row, issued by one lane each ----
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_bar_init(uint32_t bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_bar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_bar_wait(uint32_t bar) {      // phase 0: the barrier is used once per launch
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar) : "memory");
  }
}

// Everything a warp does for its env up to the warp-wide reductions of the reward terms; `sw` = this warp's shared-memory
// slice.  Lane 0's `o` is complete (clip length / loop mode / return-tracker state are loaded by lane 0 only).
__device__ __forceinline__ void env_step_fast_main(const StepParams& p, float* const sw, const uint32_t bar, const int lane, const int e,
                                                   FastTail& o) {
  using namespace fast;
  const addk_task& tk = p.task;
  float* const s_rows = sw + O_ROWS;
  float* const s_refv = sw + O_REFV;
  float* const s_sim = sw + O_SIM;
  float* const s_hist = sw + O_HIST;
  float* const s_obs = sw + O_OBS;
  float* const s_disc = sw + O_DISC;
  float* const s_demo = sw + O_DEMO;
  const bool push = (p.flags & F_UPDATE_MOTION) != 0;

  // ---- issue every independent global load first (the kernel is latency-bound: in-order issue would otherwise
  //      serialise the round trips behind the first shared-memory store that waits for its data) -----------------
  float t = 0.f;
  if (lane == 0) t = p.env.time_buf[e];
  const long long mid = p.env.motion_ids[e];
  const float off = p.env.motion_time_offsets[e];
  float sim_dp = 0.f, sim_dv = 0.f, sim_p = 0.f, sim_v = 0.f, sim_a = 0.f, sim_q = 0.f, w_dof = 0.f;
  if (lane < D) {
    sim_dp = p.sim.dof_pos[(size_t)e * p.sim.ld_dof_pos + lane];
    sim_dv = p.sim.dof_vel[(size_t)e * p.sim.ld_dof_vel + lane];
    if (p.flags & F_REWARD_DONE) w_dof = p.dof_err_w[lane];
  }
  if (lane < 3) {
    sim_p = p.sim.root_pos[(size_t)e * p.sim.ld_root_pos + lane];
    sim_v = p.sim.root_vel[(size_t)e * p.sim.ld_root_vel + lane];
    sim_a = p.sim.root_ang[(size_t)e * p.sim.ld_root_ang + lane];
  }
  if (lane < 4) sim_q = p.sim.root_rot[(size_t)e * p.sim.ld_root_rot + lane];
  float* const g_hist = p.env.hist + (size_t)e * NH * p.env.hist_stride;
  const int nhist = push ? 2 : 3;                // with a push the newest history entry is the simulator state itself
  // history rows (pose halves): one 144-byte bulk copy each, lanes 10 .. 10 + nhist - 1 (the table rows follow below:
  // 13 copy instructions per env instead of 117 float4 loads, their address arithmetic and the register -> shared stores)
  if (lane == 0) { bulk_bar_init(bar); bulk_bar_expect(bar, (uint32_t)((10 + nhist) * HALF * sizeof(float))); }
  __syncwarp();
  if (lane >= 10 && lane < 10 + nhist) {
    const int hj = lane - 10;
    const int slot = (p.newest_slot + 1 + hj) % NH;   // logical j (oldest..newest) lives in slot (newest_slot+1+j) % NH
    bulk_copy_g2s(smem_addr(s_hist + HALF * hj), g_hist + (size_t)slot * p.env.hist_stride, HALF * sizeof(float), bar);
  }
  int c_valid = 0, c_la = -1, c_lb = -1;
  if ((p.flags & F_REWARD_DONE) && lane < p.sim.contact_slots && p.sim.valid) {     // first 32 slots, loaded up front
    const size_t ci = (size_t)e * p.sim.ld_contact + lane;
    c_valid = p.sim.valid[ci]; c_la = p.sim.link_a[ci]; c_lb = p.sim.link_b[ci];
  }
  float ret0 = 0.f; long long len0 = 0;
  if ((p.flags & F_REWARD_DONE) && lane == 0 && p.env.return_buf) { ret0 = p.env.return_buf[e]; len0 = p.env.ep_len_buf[e]; }
  // ---- dependent chain: clip -> start row / length; time -> table rows -----------------------------------
  const long long start = p.lib.start_idx[mid];
  float m_len = 0.f; int m_loop = 0;
  if (lane == 0) { m_len = p.lib.lengths[mid]; m_loop = p.lib.loop_modes[mid]; }
  if (lane == 0 && (p.flags & F_ADVANCE)) {
    t = add_rn(t, tk.ctrl_dt);
    p.env.time_buf[e] = t;
  }
  t = __shfl_sync(0xffffffffu, t, 0);
  const float mt = add_rn(t, off);
  // lane b < 9 owns pose block b (0 = ref at t, 1..6 = targets, 7..8 = demo t-0.02 / t-0.01), lane 9 the velocity half of
  // the ref row: each computes its table row and issues ONE bulk copy of that half row into the staging area
  if (lane < 10) {
    float tt = mt;
    if (lane >= 1 && lane <= 6) tt = add_rn(mt, tk.tar_offsets[lane - 1]);
    if (lane >= 7 && lane <= 8) tt = add_rn(mt, tk.disc_offsets[lane - 7]);
    const long long row = table_row(p.lib, tt, tk.dt_inv, start);
    bulk_copy_g2s(smem_addr(lane < 9 ? s_rows + HALF * lane : s_refv), p.lib.table + (size_t)row * RS + (lane < 9 ? 0 : HALF),
                  HALF * sizeof(float), bar);
  }
  // ---- the simulator state into shared memory (registers -> shared), then wait for the copies ---------------------
  if (lane < D) { s_sim[7 + lane] = sim_dp; s_sim[HALF + 6 + lane] = sim_dv; }
  if (lane < 3) { s_sim[lane] = sim_p; s_sim[HALF + lane] = sim_v; s_sim[HALF + 3 + lane] = sim_a; }
  if (lane < 4) s_sim[3 + lane] = sim_q;
  if (lane == 31) { s_sim[HALF + 6 + D] = 0.0f; }          // the one pad float of the velocity half (7 + D == HALF)
  bulk_bar_wait(bar);
  __syncwarp();
  if (push) {
    if (lane < RS / 4) {
      const float4 v = *reinterpret_cast<const float4*>(s_sim + 4 * lane);
      stg4(g_hist + (size_t)p.newest_slot * p.env.hist_stride + 4 * lane, v);
      if (lane < HALF / 4) stg4(s_hist + 2 * HALF + 4 * lane, v);
    }
    // _update_ref_motion: ref_* <- table row (add_observation.py:163-175)
    if (lane < D) {
      p.env.ref_dof_pos[(size_t)e * D + lane] = s_rows[7 + lane];
      p.env.ref_dof_vel[(size_t)e * D + lane] = s_refv[6 + lane];
    }
    if (lane < 3) {
      p.env.ref_root_pos[(size_t)e * 3 + lane] = s_rows[lane];
      p.env.ref_root_vel[(size_t)e * 3 + lane] = s_refv[lane];
      p.env.ref_root_ang_vel[(size_t)e * 3 + lane] = s_refv[3 + lane];
    }
    if (lane < 4) p.env.ref_root_rot[(size_t)e * 4 + lane] = s_rows[3 + lane];
  }
  __syncwarp();

  // ---- 13 quaternions -> tangent / normal, one per lane ----------------------------------------------
  //   lane 0      simulator root            -> obs[1..6]
  //   lanes 1..6  target k = lane-1         -> obs[36 + 38k + 3 ..]
  //   lanes 7..9  history j = lane-7        -> disc[38j + 3 ..]
  //   lanes 10..12 demo j = lane-10         -> demo[38j + 3 ..]   (j = 2 is the ref row itself)
  if (lane < 13) {
    const float* src;
    float* dst;
    if (lane == 0) { src = s_sim + 3; dst = s_obs + 1; }
    else if (lane <= 6) { src = s_rows + HALF * lane + 3; dst = s_obs + 36 + STEP * (lane - 1) + 3; }
    else if (lane <= 9) { src = s_hist + HALF * (lane - 7) + 3; dst = s_disc + STEP * (lane - 7) + 3; }
    else { const int j = lane - 10; src = s_rows + HALF * (j < 2 ? 7 + j : 0) + 3; dst = s_demo + STEP * j + 3; }
    float o6[6];
    quat_to_tan_norm(ldq(src), o6);
#pragma unroll
    for (int i = 0; i < 6; ++i) dst[i] = o6[i];
  }
  // ---- copies ---------------------------------------------------------------------------------------------
  if (lane == 31) s_obs[0] = s_sim[2];                                  // root height
  if (lane < D) s_obs[7 + lane] = s_sim[7 + lane];                      // dof_pos
  if (lane < 3 * NT) {                                                  // target position: xy relative to the root, z absolute
    const int k = lane / 3, c = lane - 3 * k;
    const float v = s_rows[HALF * (1 + k) + c];
    s_obs[36 + STEP * k + c] = c < 2 ? sub_rn(v, s_sim[c]) : v;
  }
  if (lane < 3 * NH) {                                                  // discriminator positions (global obs: xyz kept)
    const int j = lane / 3, c = lane - 3 * j;
    s_disc[STEP * j + c] = s_hist[HALF * j + c];
    s_demo[STEP * j + c] = s_rows[HALF * (j < 2 ? 7 + j : 0) + c];
  }
  if (lane < D) {
#pragma unroll
    for (int k = 0; k < NT; ++k) s_obs[36 + STEP * k + 9 + lane] = s_rows[HALF * (1 + k) + 7 + lane];
#pragma unroll
    for (int j = 0; j < NH; ++j) {
      s_disc[STEP * j + 9 + lane] = s_hist[HALF * j + 7 + lane];
      s_demo[STEP * j + 9 + lane] = s_rows[HALF * (j < 2 ? 7 + j : 0) + 7 + lane];
    }
  }
  __syncwarp();

  // ---- stream the rows out ----------------------------------------------------------------------------
  {
    float* g = p.env.obs_buf + (size_t)e * OBS;
    float* gx = p.has_exp ? p.exp.next_obs + (size_t)e * OBS : nullptr;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int i = lane + 32 * r;
      if (i < OBS / 4) {
        const float4 v = *reinterpret_cast<const float4*>(s_obs + 4 * i);
        stg4(g + 4 * i, v);
        if (gx) stg4_cs(gx + 4 * i, v);
      }
    }
    float* gd = p.env.disc_obs + (size_t)e * DISC;
    float* gm = p.env.disc_obs_demo + (size_t)e * DISC;
    float* xd = p.has_exp ? p.exp.disc_obs + (size_t)e * DISC : nullptr;
    float* xm = p.has_exp ? p.exp.disc_obs_demo + (size_t)e * DISC : nullptr;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = lane + 32 * r;
      if (i < DISC / 2) {
        const float2 a = *reinterpret_cast<const float2*>(s_disc + 2 * i);
        const float2 b = *reinterpret_cast<const float2*>(s_demo + 2 * i);
        *reinterpret_cast<float2*>(gd + 2 * i) = a;
        *reinterpret_cast<float2*>(gm + 2 * i) = b;
        if (xd) { __stcs(reinterpret_cast<float2*>(xd + 2 * i), a); __stcs(reinterpret_cast<float2*>(xm + 2 * i), b); }
      }
    }
  }
  if (p.has_exp && lane == 0) {
    p.exp.motion_ids[e] = mid;
    p.exp.motion_times[e] = mt;
  }
  if (!(p.flags & F_REWARD_DONE)) return;

  // ---- tracking reward (add_reward.py:104-177) and done flags (add_done.py:97-147) --------------------------
  float pe = 0.f, ve = 0.f, de = 0.f;
  if (lane < D) {
    const float pd = sub_rn(s_rows[7 + lane], sim_dp);
    const float vd = sub_rn(s_refv[6 + lane], sim_dv);
    pe = mul_rn(mul_rn(w_dof, pd), pd);
    ve = mul_rn(mul_rn(w_dof, vd), vd);
    de = mul_rn(pd, pd);
  }
  pe = warp_sum(pe); ve = warp_sum(ve); de = warp_sum(de);
  int contact = 0;
  if (c_valid) {
    const bool ha = c_la >= 0 && c_la < 64 && ((tk.noncontact_link_mask >> c_la) & 1ull);
    const bool hb = c_lb >= 0 && c_lb < 64 && ((tk.noncontact_link_mask >> c_lb) & 1ull);
    contact = (ha || hb) ? 1 : 0;
  }
  if ((p.flags & F_REWARD_DONE) && (p.sim.contact_slots > 32 || p.sim.contact_link_masks))   // wider lists / link bitmasks
    contact |= contact_hit(p.sim, tk.noncontact_link_mask, e, lane, p.sim.contact_link_masks ? 0 : 32);
  contact = __any_sync(0xffffffffu, contact);
  o.pe = pe; o.ve = ve; o.de = de; o.contact = contact; o.t = t; o.mt = mt; o.m_len = m_len; o.m_loop = m_loop;
  o.ret0 = ret0; o.len0 = len0;
}

// The scalar end of the step: reward terms -> reward, done flag, return tracker.  ONE thread per env; `sw` = the shared
// memory slice env_step_fast_main filled for that env.
__device__ __forceinline__ void env_step_fast_tail(const StepParams& p, const int e, const float* const sw, const FastTail& o) {
  using namespace fast;
  const addk_task& tk = p.task;
  const float* const s_rows = sw + O_ROWS;
  const float* const s_refv = sw + O_REFV;
  const float* const s_sim = sw + O_SIM;
  const float pe = o.pe, ve = o.ve, de = o.de, t = o.t, mt = o.mt, m_len = o.m_len, ret0 = o.ret0;
  const int contact = o.contact, m_loop = o.m_loop;
  const long long len0 = o.len0;
  const Vec3 rp = {s_sim[0], s_sim[1], s_sim[2]}, tp = {s_rows[0], s_rows[1], s_rows[2]};
  const Vec3 dpos = {sub_rn(tp.x, rp.x), sub_rn(tp.y, rp.y), sub_rn(tp.z, rp.z)};
  Vec3 dpos_r = dpos;
  if (!tk.track_root) { dpos_r.x = 0.f; dpos_r.y = 0.f; }
  if (!tk.track_root_h) dpos_r.z = 0.f;
  const float root_pos_err = add_rn(add_rn(mul_rn(dpos_r.x, dpos_r.x), mul_rn(dpos_r.y, dpos_r.y)), mul_rn(dpos_r.z, dpos_r.z));
  Quat rq = ldq(s_sim + 3), tq = ldq(s_rows + 3);
  Vec3 rv = {s_sim[HALF], s_sim[HALF + 1], s_sim[HALF + 2]}, ra = {s_sim[HALF + 3], s_sim[HALF + 4], s_sim[HALF + 5]};
  Vec3 tv = {s_refv[0], s_refv[1], s_refv[2]}, ta = {s_refv[3], s_refv[4], s_refv[5]};
  if (!tk.track_root) {  // convert_to_local_root (add_reward.py:91-102)
    const Quat h0 = calc_heading_quat_inv(rq), h1 = calc_heading_quat_inv(tq);
    rv = quat_rotate(h0, rv); ra = quat_rotate(h0, ra); rq = quat_mul(h0, rq);
    tv = quat_rotate(h1, tv); ta = quat_rotate(h1, ta); tq = quat_mul(h1, tq);
  }
  float rot_err = quat_diff_angle(rq, tq);
  rot_err = mul_rn(rot_err, rot_err);
  const Vec3 dv = {sub_rn(tv.x, rv.x), sub_rn(tv.y, rv.y), sub_rn(tv.z, rv.z)};
  const Vec3 da = {sub_rn(ta.x, ra.x), sub_rn(ta.y, ra.y), sub_rn(ta.z, ra.z)};
  const float vel_err = add_rn(add_rn(mul_rn(dv.x, dv.x), mul_rn(dv.y, dv.y)), mul_rn(dv.z, dv.z));
  const float ang_err = add_rn(add_rn(mul_rn(da.x, da.x), mul_rn(da.y, da.y)), mul_rn(da.z, da.z));
  const float pose_r = expf(mul_rn(-tk.pose_scale, pe));
  const float vel_r = expf(mul_rn(-tk.vel_scale, ve));
  const float root_pose_r = expf(mul_rn(-tk.root_pose_scale, add_rn(root_pos_err, mul_rn(0.1f, rot_err))));
  const float root_vel_r = expf(mul_rn(-tk.root_vel_scale, add_rn(vel_err, mul_rn(0.1f, ang_err))));
  const float r = add_rn(add_rn(add_rn(mul_rn(tk.pose_w, pose_r), mul_rn(tk.vel_w, vel_r)), mul_rn(tk.root_pose_w, root_pose_r)),
                         mul_rn(tk.root_vel_w, root_vel_r));
  int done = 0;
  if (t >= tk.ep_len) done = 3;
  if (mt >= m_len && m_loop != 1) done = 2;
  if (tk.enable_early_termination) {
    bool failed = contact != 0;
    if (tk.pose_termination) {
      bool pf = (de / (float)D) > tk.pose_termination_dist;
      if (tk.track_root) {
        const float re = add_rn(add_rn(mul_rn(dpos.x, dpos.x), mul_rn(dpos.y, dpos.y)), mul_rn(dpos.z, dpos.z));
        pf = pf || (re > tk.pose_termination_dist);
      }
      failed = failed || pf;
    }
    if (failed && t > 0.0f) done = 1;
  }
  p.env.reward[e] = r;
  p.env.done[e] = done;
  if (p.has_exp) { p.exp.reward[e] = r; p.exp.done[e] = done; }
  if (p.env.return_buf) {        // ReturnTracker.update (base_agent.py:596-621)
    float ret = add_rn(ret0, r);
    long long len = len0 + 1;
    if (done != 0) {
      atomicAdd(p.env.tracker_sums, (double)ret);
      atomicAdd(p.env.tracker_sums + 1, (double)len);
      atomicAdd(reinterpret_cast<unsigned long long*>(p.env.tracker_count), 1ull);
      p.env.eps_per_env[e] += 1;
      ret = 0.f; len = 0;
    }
    p.env.return_buf[e] = ret;
    p.env.ep_len_buf[e] = len;
  }
}

// The warps leave their reduced reward terms in shared memory and lanes 0..WPB-1 of warp 0 run the scalar end for the
// block's envs at once (lane 0 of every warp running it for its own env was 273 warp-instructions of one-lane work per
// env: a quarter of the kernel's issue slots, ncu r01_step_fast).  Register caps for more resident blocks (48 / 40
// registers: 40 / 128 B of spills) measured slower and are gone (profiles/r01_SUMMARY.md).
__global__ void __launch_bounds__(WPB * 32) env_step_fast_kernel(const __grid_constant__ StepParams p) {
  using namespace fast;
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * WPB + warp;
  const bool active = e < p.n && !((p.flags & F_MASKED) && !p.env_mask[e]);
  float* const sw = smem + (size_t)warp * PER_WARP;
  __shared__ __align__(8) unsigned long long s_bar[WPB];      // one single-use mbarrier per warp (bulk-copy staging)
  const uint32_t bar = smem_addr(&s_bar[warp]);
  FastTail o;
  __shared__ FastTail s_tail[WPB];
  __shared__ int s_active[WPB];
  if (active) env_step_fast_main(p, sw, bar, lane, e, o);
  if (!(p.flags & F_REWARD_DONE)) return;                 // uniform over the grid
  if (lane == 0) { s_active[warp] = active ? 1 : 0; if (active) s_tail[warp] = o; }
  __syncthreads();
  if (warp == 0 && lane < WPB && s_active[lane])
    env_step_fast_tail(p, blockIdx.x * WPB + lane, smem + (size_t)lane * PER_WARP, s_tail[lane]);
}

// ------------------------------------------------------------------------------------------------
struct ResetParams {
  addk_task task;
  addk_motion_lib lib;
  addk_env_buffers env;
  const long long* new_ids;
  const float* new_times;
  int n, head, reset_all, zero_time_done;
  float* qpos_out;
  float* qvel_out;
  uint8_t* reset_mask_out;
};

__global__ void __launch_bounds__(WPB * 32) reset_done_kernel(const __grid_constant__ ResetParams p) {
  const addk_task& tk = p.task;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * WPB + warp;
  if (e >= p.n) return;
  int dn = (lane == 0) ? p.env.done[e] : 0;   // lane 0 also clears the flag below
  dn = __shfl_sync(0xffffffffu, dn, 0);
  const bool doit = p.reset_all || dn != 0;
  if (lane == 0 && p.reset_mask_out) p.reset_mask_out[e] = doit ? 1 : 0;
  if (!doit) return;
  const int D = tk.num_dofs, half = (7 + D + 3) & ~3, RS = p.lib.row_stride, nH = tk.num_disc_steps;
  const long long mid = p.new_ids[e];
  const float off = p.new_times[e];
  if (lane == 0) {
    p.env.motion_ids[e] = mid;
    p.env.motion_time_offsets[e] = off;
    if (p.zero_time_done) {
      p.env.time_buf[e] = 0.0f;   // Environment.reset_idx (env.py:157-163)
      p.env.done[e] = 0;          // ADDDone.reset_idx (add_done.py:92-93)
    }
  }
  const float mt = add_rn(p.zero_time_done ? 0.0f : p.env.time_buf[e], off);
  const long long start = p.lib.start_idx[mid];
  const float* row = p.lib.table + (size_t)table_row(p.lib, mt, tk.dt_inv, start) * RS;
  // ref_* and the pose written into the simulator (add_observation.py:308-331)
  for (int c = lane; c < 7 + D; c += 32) {
    float v = row[c];
    if (p.qpos_out) p.qpos_out[(size_t)e * (7 + D) + c] = v;
    if (c < 3) p.env.ref_root_pos[(size_t)e * 3 + c] = v;
    else if (c < 7) p.env.ref_root_rot[(size_t)e * 4 + (c - 3)] = v;
    else p.env.ref_dof_pos[(size_t)e * D + (c - 7)] = v;
  }
  for (int c = lane; c < 6 + D; c += 32) {
    float v = row[half + c];
    if (p.qvel_out) p.qvel_out[(size_t)e * (6 + D) + c] = v;
    if (c < 3) p.env.ref_root_vel[(size_t)e * 3 + c] = v;
    else if (c < 6) p.env.ref_root_ang_vel[(size_t)e * 3 + (c - 3)] = v;
    else p.env.ref_dof_vel[(size_t)e * D + (c - 6)] = v;
  }
  // _reset_disc_hist + CircularBuffer.fill: logical j -> slot (head + j) % nH (circular_buffer.py:22-29)
  float* g_hist = p.env.hist + (size_t)e * nH * p.env.hist_stride;
  const int r4 = RS / 4;
  for (int i = lane; i < nH * r4; i += 32) {
    int j = i / r4, c = i - j * r4;
    const float* src = p.lib.table + (size_t)table_row(p.lib, add_rn(mt, tk.disc_offsets[j]), tk.dt_inv, start) * RS;
    int slot = (p.head + j) % nH;
    stg4(g_hist + (size_t)slot * p.env.hist_stride + 4 * c, ldg4(src + 4 * c));
  }
}

// ------------------------------------------------------------------------------------------------
// c10::div_floor_floating, the arithmetic behind `time // dt` on fp32 tensors (sampler.py:88).
__device__ __forceinline__ float floor_div(float a, float b) {
  float mod = fmodf(a, b);
  float div = sub_rn(a, mod) / b;
  if (mod != 0.0f && ((b < 0.0f) != (mod < 0.0f))) div = sub_rn(div, 1.0f);
  float fl;
  if (div != 0.0f) {
    fl = floorf(div);
    if (sub_rn(div, fl) > 0.5f) fl = add_rn(fl, 1.0f);
  } else {
    fl = copysignf(0.0f, a / b);
  }
  return fl;
}

__global__ void sample_clip_kernel(const float* __restrict__ weights, int C, const float* __restrict__ errors, int S,
                                   const int32_t* __restrict__ done, const float* __restrict__ u, int n,
                                   unsigned int* temp_bits, long long* ids_out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  if (done && done[e] == 0) return;
  // inverse CDF over the (normalised) clip weights -- torch.multinomial with replacement (motion_lib.py:35-39)
  // (the fp32 running sum can end below 1: the fallback is the last clip with a positive weight, never a
  // zero-weight clip -- torch.multinomial cannot draw those either)
  float u0 = u[3 * (size_t)e], cum = 0.f;
  int c = -1, last_pos = 0;
  for (int k = 0; k < C; ++k) {
    const float w = weights[k];
    if (w > 0.f) last_pos = k;
    cum = add_rn(cum, w);
    if (u0 < cum && w > 0.f) { c = k; break; }
  }
  if (c < 0) c = last_pos;
  ids_out[e] = c;
  if (temp_bits) {  // temperature None: max error over the clips drawn in this batch (sampler.py:66-69)
    float m = 0.f;
    for (int s = 0; s < S; ++s) m = fmaxf(m, errors[c * S + s]);
    atomicMax(temp_bits, __float_as_uint(m));
  }
}

// AdaptiveSegmentSampler.sample_start_frame after the draws (sampler.py:86-92): seg * size + U * size, (t // dt) * dt,
// clamp(min = min_start) -- every product and sum rounded to fp32 on its own, like the reference's tensor ops.
__device__ __forceinline__ float start_time_of(int seg, float u2, float sz, float dt, float min_start) {
  float tm = add_rn(mul_rn((float)seg, sz), mul_rn(u2, sz));
  tm = mul_rn(floor_div(tm, dt), dt);
  return fmaxf(tm, min_start);
}

// the same arithmetic from GIVEN (segment, U) draws: what a host that keeps torch.multinomial / torch.rand calls
__global__ void start_time_from_draws_kernel(const float* __restrict__ seg_sizes, float dt, float min_start,
                                             const long long* __restrict__ clip_ids, const long long* __restrict__ segs,
                                             const float* __restrict__ u, int n, float* __restrict__ times_out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  times_out[e] = start_time_of((int)segs[e], u[e], seg_sizes[clip_ids[e]], dt, min_start);
}

__global__ void sample_time_kernel(const float* __restrict__ errors, int S, const float* __restrict__ seg_sizes,
                                   float dt, float min_start, float temperature, int rand_reset,
                                   const int32_t* __restrict__ done, const float* __restrict__ u, int n,
                                   const unsigned int* temp_bits, const long long* __restrict__ ids, float* times_out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  if (done && done[e] == 0) return;
  if (!rand_reset) { times_out[e] = 0.f; return; }
  int c = (int)ids[e];
  float temp = temp_bits ? add_rn(__uint_as_float(*temp_bits), 1e-6f) : temperature;
  const float* er = errors + (size_t)c * S;
  float mx = -INFINITY;
  for (int s = 0; s < S; ++s) mx = fmaxf(mx, er[s] / temp);
  float z = 0.f;
  for (int s = 0; s < S; ++s) z = add_rn(z, expf(sub_rn(er[s] / temp, mx)));
  float u1 = u[3 * (size_t)e + 1], cum = 0.f;
  int seg = S - 1;
  for (int s = 0; s < S; ++s) { cum = add_rn(cum, expf(sub_rn(er[s] / temp, mx)) / z); if (u1 < cum) { seg = s; break; } }
  times_out[e] = start_time_of(seg, u[3 * (size_t)e + 2], seg_sizes[c], dt, min_start);
}

// One warp per sample; the (clip, segment) bins are first accumulated per BLOCK in shared memory (131072 double atomics on
// the 20 bins of a single clip took 241 us), then every block adds its non-empty bins to the global sums.  Libraries with
// more bins than SAMPLER_SMEM_BINS fall back to direct global atomics.
constexpr int SAMPLER_SMEM_BINS = 1024;
__global__ void sampler_accum_kernel(const long long* __restrict__ clip_ids, const float* __restrict__ timesteps,
                                     const float* __restrict__ a, const float* __restrict__ b, int dim, int n,
                                     const float* __restrict__ seg_sizes, int S, int bins, double* sums, int* counts) {
  __shared__ double s_sum[SAMPLER_SMEM_BINS];
  __shared__ int s_cnt[SAMPLER_SMEM_BINS];
  const bool local = bins <= SAMPLER_SMEM_BINS;
  if (local) {
    for (int i = threadIdx.x; i < bins; i += blockDim.x) { s_sum[i] = 0.0; s_cnt[i] = 0; }
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32); i < n; i += warps) {
    float acc = 0.f;
    if (b) {  // tracking error = sum((disc_obs - disc_obs_demo)^2)  (add_agent.py:120-122)
      for (int c = lane; c < dim; c += 32) {
        float d = sub_rn(a[(size_t)i * dim + c], b[(size_t)i * dim + c]);
        acc += d * d;
      }
      acc = warp_sum(acc);
    } else {  // `a` already holds one tracking error per sample
      acc = a[i];
    }
    if (lane == 0) {
      long long c = clip_ids[i];
      float sz = fmaxf(seg_sizes[c], 1e-6f);
      long long seg = (long long)(timesteps[i] / sz);
      seg = seg < 0 ? 0 : (seg > S - 1 ? S - 1 : seg);
      const long long bin = c * S + seg;
      if (local) { atomicAdd(&s_sum[bin], (double)acc); atomicAdd(&s_cnt[bin], 1); }
      else { atomicAdd(sums + bin, (double)acc); atomicAdd(counts + bin, 1); }
    }
  }
  if (local) {
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += blockDim.x)
      if (s_cnt[i]) { atomicAdd(sums + i, s_sum[i]); atomicAdd(counts + i, s_cnt[i]); }
  }
}

__global__ void sampler_ema_kernel(double* sums, int* counts, int bins, float* errors) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bins) return;
  if (counts[i] > 0) {
    float mean = (float)(sums[i] / (double)counts[i]);
    errors[i] = add_rn(mul_rn(0.9f, errors[i]), mul_rn(0.1f, mean));
  }
  sums[i] = 0.0;
  counts[i] = 0;
}

}  // namespace addk

using namespace addk;

static int step_smem_bytes(const addk_task& tk, int RS) {
  const int D = tk.num_dofs, half = (7 + D + 3) & ~3;
  const int nT = tk.enable_tar_obs ? tk.num_tar_steps : 0, nH = tk.num_disc_steps;
  const int per_warp = RS + nT * half + nH * RS + RS + nH * RS + ((tk.obs_dim + 3) & ~3) + 2 * ((tk.disc_obs_dim + 3) & ~3);
  return per_warp * WPB * (int)sizeof(float);
}

extern "C" int addk_env_step(void* stream, const addk_task* task, const addk_motion_lib* lib,
                             const addk_sim_state* sim, const addk_env_buffers* env, const addk_exp_row* exp,
                             const float* dof_err_w, const uint8_t* env_mask, int num_envs, int hist_head, int flags) {
  if (!task || !lib || !sim || !env || num_envs <= 0) return ADDK_ERR_ARG;
  const int D = task->num_dofs, half = (7 + D + 3) & ~3;
  if (D < 1 || D > 31 || lib->row_stride < 2 * half || (lib->row_stride & 3) || env->hist_stride < lib->row_stride ||
      (env->hist_stride & 3) || task->num_tar_steps > ADDK_MAX_TAR_STEPS || task->num_tar_steps > 16 ||
      task->num_disc_steps > ADDK_MAX_DISC_STEPS || task->num_disc_steps < 1 || sim->contact_slots < 0 ||
      (sim->contact_slots > 0 && sim->valid && (!sim->link_a || !sim->link_b || sim->ld_contact < sim->contact_slots)))
    return ADDK_ERR_ARG;
  if ((flags & F_MASKED) && !env_mask) return ADDK_ERR_ARG;
  if ((flags & F_REWARD_DONE) && !dof_err_w) return ADDK_ERR_ARG;
  StepParams p;
  p.task = *task; p.lib = *lib; p.sim = *sim; p.env = *env;
  p.has_exp = exp != nullptr;
  if (exp) p.exp = *exp; else p.exp = addk_exp_row{};
  p.dof_err_w = dof_err_w; p.env_mask = env_mask; p.n = num_envs; p.flags = flags;
  const int nH = task->num_disc_steps;
  // with a push the newest entry lands in slot `hist_head`; without, the newest is the slot before the head
  p.newest_slot = (flags & F_UPDATE_MOTION) ? hist_head % nH : (hist_head + nH - 1) % nH;
  int smem = step_smem_bytes(*task, lib->row_stride);
  if (smem > 200 * 1024) return ADDK_ERR_UNSUPPORTED;
  const bool fast = D == 29 && lib->row_stride == 72 && task->enable_tar_obs && task->num_tar_steps == 6 &&
                    task->num_disc_steps == 3 && task->global_obs && task->root_height_obs && !task->enable_vel_obs &&
                    !task->enable_phase_obs && task->obs_dim == 264 && task->disc_obs_dim == 114 &&
                    task->disc_offsets[2] == 0.0f;   // the newest demo row is the reference row itself
  static int configured = 0;
  if (smem > 48 * 1024 && configured < smem) {
    cudaFuncSetAttribute(env_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    configured = smem;
  }
  if (fast && env->hist_stride == 72) {
    const int fsm = fast::PER_WARP * WPB * (int)sizeof(float);
    env_step_fast_kernel<<<(num_envs + WPB - 1) / WPB, WPB * 32, fsm, (cudaStream_t)stream>>>(p);
  } else {
    env_step_kernel<false><<<(num_envs + WPB - 1) / WPB, WPB * 32, smem, (cudaStream_t)stream>>>(p);
  }
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}

extern "C" int addk_reset_done(void* stream, const addk_task* task, const addk_motion_lib* lib,
                               const addk_env_buffers* env, const long long* new_ids, const float* new_times,
                               int num_envs, int hist_head, int reset_all, int zero_time_done, float* qpos_out, float* qvel_out,
                               uint8_t* reset_mask_out) {
  if (!task || !lib || !env || !new_ids || !new_times || num_envs <= 0) return ADDK_ERR_ARG;
  ResetParams p;
  p.task = *task; p.lib = *lib; p.env = *env; p.new_ids = new_ids; p.new_times = new_times;
  p.n = num_envs; p.head = hist_head % task->num_disc_steps; p.reset_all = reset_all; p.zero_time_done = zero_time_done;
  p.qpos_out = qpos_out; p.qvel_out = qvel_out; p.reset_mask_out = reset_mask_out;
  reset_done_kernel<<<(num_envs + WPB - 1) / WPB, WPB * 32, 0, (cudaStream_t)stream>>>(p);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}

extern "C" int addk_sample_motion_time(void* stream, const float* motion_weights, int num_motions,
                                       const float* errors, int num_segments, const float* seg_sizes, float dt,
                                       float min_start_time, float temperature, int rand_reset,
                                       const int32_t* done, const float* uniforms, int num_envs,
                                       unsigned int* temp_bits_work, long long* ids_out, float* times_out) {
  if (!motion_weights || !errors || !seg_sizes || !uniforms || !ids_out || !times_out || num_envs <= 0)
    return ADDK_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned int* tb = (temperature < 0.f) ? temp_bits_work : nullptr;
  if (temperature < 0.f && !temp_bits_work) return ADDK_ERR_ARG;
  if (tb) cudaMemsetAsync(tb, 0, sizeof(unsigned int), st);
  int th = 128, bl = (num_envs + th - 1) / th;
  sample_clip_kernel<<<bl, th, 0, st>>>(motion_weights, num_motions, errors, num_segments, done, uniforms, num_envs,
                                        tb, ids_out);
  ADDK_CHECK_LAUNCH();
  sample_time_kernel<<<bl, th, 0, st>>>(errors, num_segments, seg_sizes, dt, min_start_time, temperature, rand_reset,
                                        done, uniforms, num_envs, tb, ids_out, times_out);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}

extern "C" int addk_sample_start_time(void* stream, const float* errors, int num_segments, const float* seg_sizes,
                                      float dt, float min_start_time, float temperature, const float* uniforms, int n,
                                      const long long* clip_ids, float* times_out) {
  if (!errors || !seg_sizes || !uniforms || !clip_ids || !times_out || n <= 0 || temperature <= 0.f) return ADDK_ERR_ARG;
  sample_time_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(errors, num_segments, seg_sizes, dt,
                                                                        min_start_time, temperature, 1, nullptr, uniforms,
                                                                        n, nullptr, clip_ids, times_out);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}

extern "C" int addk_start_time_from_draws(void* stream, const float* seg_sizes, float dt, float min_start_time,
                                          const long long* clip_ids, const long long* segments, const float* uniforms,
                                          int n, float* times_out) {
  if (!seg_sizes || !clip_ids || !segments || !uniforms || !times_out || n <= 0) return ADDK_ERR_ARG;
  start_time_from_draws_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(seg_sizes, dt, min_start_time, clip_ids,
                                                                                  segments, uniforms, n, times_out);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}

extern "C" int addk_sampler_update_errors(void* stream, const long long* clip_ids, const float* timesteps,
                                          const float* disc_obs, const float* disc_obs_demo, int disc_dim, int n,
                                          const float* seg_sizes, int num_motions, int num_segments,
                                          double* sums_work, int* counts_work, float* errors) {
  // disc_obs_demo == NULL: `disc_obs` is a ready-made [n] vector of tracking errors (sampler.py:21 signature)
  if (!clip_ids || !timesteps || !disc_obs || !sums_work || !counts_work || !errors || n <= 0) return ADDK_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (n + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  sampler_accum_kernel<<<blocks, 256, 0, st>>>(clip_ids, timesteps, disc_obs, disc_obs_demo, disc_dim, n, seg_sizes,
                                               num_segments, num_motions * num_segments, sums_work, counts_work);
  ADDK_CHECK_LAUNCH();
  int bins = num_motions * num_segments;
  sampler_ema_kernel<<<(bins + 127) / 128, 128, 0, st>>>(sums_work, counts_work, bins, errors);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}
