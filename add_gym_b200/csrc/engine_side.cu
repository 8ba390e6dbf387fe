// Engine-side hand-off: what sits between the physics backend's flat arrays and the fused step
// kernel.  The reference does this work in Python per call -- and for MuJoCo-Warp partly on the HOST:
//   * MJWarpEntity.get_contacts (engine/mjwarp_engine.py:896-986): `nacon` read back with .item(), the contact arrays copied
//     to the host, a Python loop over every contact, padded [nworld, max_len] tensors rebuilt per step;
//   * MJWarpEntity.get_pos/get_quat/get_vel/get_ang/get_dofs_position/get_dofs_velocity (mjwarp_engine.py:640-795,
//     robot.py:271-293): six getters, a fresh zeros() + a Python loop over joint segments per call;
//   * MJWarpScene.step's PD prologue (mjwarp_engine.py:1565-1604): two .item() syncs + ~10 torch ops per substep.
// Here each is ONE launch on the backend's device arrays, no host round trip; nothing needs MuJoCo to be tested.
#include "common.cuh"
#include "addk.h"

namespace addk {

// One thread per contact slot i < *nacon.  geom pair -> body pair (geom_bodyid), the reference's filters in its order
// (negative geoms skipped; self-contact excluded when self is other; (self, other) or (other, self) orientation), then
// the self-side body sets bit b in out[2 w] and the other-side body in out[2 w + 1]: the per-world link bitmasks that
// `isin(link_a, ids) & valid` / `isin(link_b, ids) & valid` (robot.py:221-231) are evaluated against.
__global__ void contact_link_mask_kernel(const int* __restrict__ geom_pairs, const int* __restrict__ world_ids,
                                         const int* __restrict__ nacon_dev, int nacon_cap, const int* __restrict__ geom_bodyid,
                                         int ngeom, unsigned long long self_mask, unsigned long long other_mask,
                                         int self_is_other, int exclude_self, int nworld, unsigned long long* __restrict__ out) {
  const int n = min(*nacon_dev, nacon_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int w = world_ids[i];
    const int g0 = geom_pairs[2 * i], g1 = geom_pairs[2 * i + 1];
    if (g0 < 0 || g1 < 0 || g0 >= ngeom || g1 >= ngeom || w < 0 || w >= nworld) continue;
    const int b0 = geom_bodyid[g0], b1 = geom_bodyid[g1];
    const bool a_self = b0 >= 0 && b0 < 64 && ((self_mask >> b0) & 1ull), b_self = b1 >= 0 && b1 < 64 && ((self_mask >> b1) & 1ull);
    if (exclude_self && self_is_other && a_self && b_self) continue;
    const bool a_other = b0 >= 0 && b0 < 64 && ((other_mask >> b0) & 1ull), b_other = b1 >= 0 && b1 < 64 && ((other_mask >> b1) & 1ull);
    if (a_self && b_other) {
      atomicOr(out + 2 * (size_t)w, 1ull << b0);
      atomicOr(out + 2 * (size_t)w + 1, 1ull << b1);
    } else if (b_self && a_other) {
      atomicOr(out + 2 * (size_t)w, 1ull << b1);
      atomicOr(out + 2 * (size_t)w + 1, 1ull << b0);
    }
  }
}

// One warp per world: gathers the 7 + D pose columns of qpos and the 6 + D velocity columns of qvel (column maps = the
// entity's segment table, BFS dof order) into the packed [pose half | velocity half] row the step kernel stages anyway.
__global__ void pack_state_kernel(const float* __restrict__ qpos, int ld_qpos, const float* __restrict__ qvel, int ld_qvel,
                                  const int* __restrict__ qpos_col, const int* __restrict__ qvel_col, int D, int nworld,
                                  float* __restrict__ out, int row_stride) {
  const int w = blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
  const int lane = threadIdx.x & 31;
  if (w >= nworld) return;
  const int half = (7 + D + 3) & ~3;
  float* row = out + (size_t)w * row_stride;
  for (int c = lane; c < half; c += 32) {
    float v = 0.f;
    if (c < 7 + D) { const int col = qpos_col[c]; v = col >= 0 ? qpos[(size_t)w * ld_qpos + col] : 0.f; }
    row[c] = v;
  }
  for (int c = lane; c < row_stride - half; c += 32) {
    float v = 0.f;
    if (c < 6 + D) { const int col = qvel_col[c]; v = col >= 0 ? qvel[(size_t)w * ld_qvel + col] : 0.f; }
    row[half + c] = v;
  }
}

// tau = kp (target - pos) - kv vel, clamped to +-max_torque, added to qfrc_applied for every local dof >= 6 whose gains
// are not both zero (mjwarp_engine.py:1574-1602).  One thread per (world, local dof); qfrc must be zeroed by the caller
// (the reference zeroes it at the top of every substep).
__global__ void pd_control_kernel(const float* __restrict__ qpos, int ld_qpos, const float* __restrict__ qvel, int ld_qvel,
                                  const float* __restrict__ target, const float* __restrict__ kp, const float* __restrict__ kv,
                                  const int* __restrict__ pos_col, const int* __restrict__ dof_ids, int n_local,
                                  float max_torque, int nworld, float* __restrict__ qfrc, int ld_qfrc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)nworld * n_local) return;
  const int w = (int)(i / n_local), j = (int)(i - (long long)w * n_local);
  if (j < 6) return;                                                  // never on the floating base
  const float p = kp[j], d = kv[j];
  if (p == 0.f && d == 0.f) return;
  const int pc = pos_col[j], dv = dof_ids[j];
  const float pos = pc >= 0 ? qpos[(size_t)w * ld_qpos + pc] : 0.f;   // ball joints report 0 (get_dofs_position)
  const float vel = qvel[(size_t)w * ld_qvel + dv];
  float tau = sub_rn(mul_rn(p, sub_rn(target[(size_t)w * n_local + j], pos)), mul_rn(d, vel));
  if (max_torque > 0.f) tau = fminf(fmaxf(tau, -max_torque), max_torque);
  qfrc[(size_t)w * ld_qfrc + dv] += tau;
}

}  // namespace addk

using namespace addk;

extern "C" int addk_contact_link_mask(void* stream, const int* geom_pairs, const int* world_ids, const int* nacon_dev,
                                      int nacon_cap, const int* geom_bodyid, int ngeom, unsigned long long self_body_mask,
                                      unsigned long long other_body_mask, int self_is_other, int exclude_self_contact,
                                      int nworld, unsigned long long* link_masks_out) {
  if (!geom_pairs || !world_ids || !nacon_dev || !geom_bodyid || !link_masks_out || nworld <= 0 || nacon_cap < 0 || ngeom <= 0)
    return ADDK_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(link_masks_out, 0, 2 * sizeof(unsigned long long) * (size_t)nworld, st) != cudaSuccess) {
    addk_set_error("contact_link_mask: memset failed"); return ADDK_ERR_LAUNCH;
  }
  if (nacon_cap == 0) return ADDK_OK;
  int blocks = (nacon_cap + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  contact_link_mask_kernel<<<blocks, 256, 0, st>>>(geom_pairs, world_ids, nacon_dev, nacon_cap, geom_bodyid, ngeom, self_body_mask,
                                                   other_body_mask, self_is_other, exclude_self_contact, nworld, link_masks_out);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}

extern "C" int addk_pack_state(void* stream, const float* qpos, int ld_qpos, const float* qvel, int ld_qvel,
                               const int* qpos_col, const int* qvel_col, int num_dofs, int nworld, float* rows_out,
                               int row_stride) {
  const int half = (7 + num_dofs + 3) & ~3;
  if (!qpos || !qvel || !qpos_col || !qvel_col || !rows_out || nworld <= 0 || num_dofs < 1 || num_dofs > 31 ||
      row_stride < half + 6 + num_dofs)
    return ADDK_ERR_ARG;
  pack_state_kernel<<<(nworld + 7) / 8, 256, 0, (cudaStream_t)stream>>>(qpos, ld_qpos, qvel, ld_qvel, qpos_col, qvel_col, num_dofs,
                                                                      nworld, rows_out, row_stride);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}

extern "C" int addk_pd_control(void* stream, const float* qpos, int ld_qpos, const float* qvel, int ld_qvel,
                               const float* target, const float* kp, const float* kv, const int* pos_col, const int* dof_ids,
                               int n_local, float max_torque, int nworld, float* qfrc, int ld_qfrc) {
  if (!qpos || !qvel || !target || !kp || !kv || !pos_col || !dof_ids || !qfrc || n_local <= 0 || nworld <= 0) return ADDK_ERR_ARG;
  const long long n = (long long)nworld * n_local;
  pd_control_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(qpos, ld_qpos, qvel, ld_qvel, target, kp, kv, pos_col,
                                                                               dof_ids, n_local, max_torque, nworld, qfrc, ld_qfrc);
  ADDK_CHECK_LAUNCH();
  return ADDK_OK;
}
