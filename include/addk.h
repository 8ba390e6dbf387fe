/* addk -- C ABI of the B200-native add-gym hot path (libaddk.so).
 *
 * The reference (rsamf/add-gym) has no FFI: its boundary for this path is a set of Python classes the
 * agent constructs by name (add_gym/learning/add/add_agent.py:30-60).  The drop-in classes in
 * add_gym_b200/ keep those names and put this library underneath them.  Every entry point below cites
 * the reference interface it replaces; INTEGRATION.md shows the ctypes stub a maintainer would add to
 * the reference itself.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`; the caller owns all memory,
 *     nothing is allocated and nothing synchronises with the host inside a call;
 *   - `stream` is a cudaStream_t (pass torch.cuda.current_stream().cuda_stream);
 *   - return value 0 = ok, non-zero = ADDK_ERR_* ; addk_last_error() gives the text;
 *   - quaternions are wxyz, all floating data fp32, ids int64 (torch.long), done flags int32.
 */
#ifndef ADDK_H_
#define ADDK_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADDK_MAX_TAR_STEPS 16
#define ADDK_MAX_DISC_STEPS 8

const char* addk_last_error(void);
int addk_version(void);
/* number of kernel launches issued through this library since the last reset (bench.py "gpu_launches") */
long long addk_launch_count(int reset);
/* bit 0: built with the superseded tf32 / tf32x3 kernels (make LEGACY=1; precision modes 1 and 2 need it) */
int addk_build_flags(void);
/* test / profiling hooks: id of the kernel the last addk_gemm call was dispatched to (0 CUDA-core fp32, 10 tf32,
 * 11 tf32x3, 12 tf32x3 CTA pair, 20 bf16 one tile per CTA, 21 bf16 persistent, 22 bf16 persistent CTA pairs, 30 f16x3 one
 * tile per CTA, 31 f16x3 persistent, 32 f16x3 persistent CTA pairs (cta_group::2)); a device buffer of 16 int64 that CTA 0 of the persistent kernels fills with clock stamps
 * (NULL switches the stamps off, the default) */
int addk_debug_last_gemm_kernel(void);
int addk_debug_set_stamp_buffer(long long* device_buffer16);

/* ---------------------------------------------------------------------------------------------
 * Task description shared by the per-step kernels.  Plain data, filled once by the host from the
 * task YAML keys the reference plugins read (configs/task/pose.yaml).
 * ------------------------------------------------------------------------------------------- */
typedef struct addk_task {
  int32_t num_dofs;            /* D = 29 */
  int32_t num_tar_steps;       /* len(tar_obs_steps); 0 when enable_tar_obs is false */
  int32_t num_disc_steps;      /* num_disc_obs_steps */
  int32_t global_obs, root_height_obs, enable_vel_obs, enable_phase_obs, enable_tar_obs;
  int32_t num_phase_encoding;
  int32_t obs_dim, disc_obs_dim;
  int32_t track_root, track_root_h;                 /* add_obs._track_global_root(), root_height_obs */
  int32_t enable_early_termination, pose_termination;
  int32_t contact_slots;                            /* unused since round 2: the width travels per step in addk_sim_state */
  float tar_offsets[ADDK_MAX_TAR_STEPS];            /* fl32(dt * k)  (add_observation.py:214-215) */
  float disc_offsets[ADDK_MAX_DISC_STEPS];          /* oldest -> newest (add_observation.py:362-370) */
  float ctrl_dt;                                    /* env.time_buf += ctrl_dt (env.py:155) */
  float dt_inv;                                     /* MotionLib._dt_inv = round(1/dt) (motion_lib.py:23) */
  float pose_w, vel_w, root_pose_w, root_vel_w;     /* add_reward.py:12-21 */
  float pose_scale, vel_scale, root_pose_scale, root_vel_scale;
  float ep_len, pose_termination_dist;              /* add_done.py:21-25 */
  uint64_t noncontact_link_mask;                    /* bit l set: link idx l is not allowed to touch the plane */
} addk_task;

/* Motion library metadata living on the device (MotionLib, motion_lib.py:236-283). */
typedef struct addk_motion_lib {
  const float* table;          /* [S_total, row_stride] step table, layout in csrc/motion.cu */
  int32_t row_stride;          /* floats per row (72 for D = 29) */
  int32_t num_motions;
  int64_t s_total;             /* rows in the table = _step_root_pos.shape[-2] */
  const int64_t* start_idx;    /* [C]  MotionLib._motion_start_idx (30 fps cumsum, quirk Q2) */
  const float* lengths;        /* [C]  seconds */
  const int32_t* loop_modes;   /* [C]  0 CLAMP / 1 WRAP */
} addk_motion_lib;

/* Simulator state as the engine getters return it (robot.py:271-293): AoS rows with a leading dim. */
typedef struct addk_sim_state {
  const float* root_pos;  int32_t ld_root_pos;     /* [N,3]   */
  const float* root_rot;  int32_t ld_root_rot;     /* [N,4]   */
  const float* root_vel;  int32_t ld_root_vel;     /* [N,3]   */
  const float* root_ang;  int32_t ld_root_ang;     /* [N,3]   */
  const float* dof_pos;   int32_t ld_dof_pos;      /* [N,D] = get_dofs_position()[:,6:] */
  const float* dof_vel;   int32_t ld_dof_vel;      /* [N,D] */
  const int32_t* link_a;  const int32_t* link_b;   /* [N,C] contact pairs vs the ground plane */
  const uint8_t* valid;                            /* [N,C]; NULL = no contact list this step */
  int32_t contact_slots;                           /* C = get_contacts()[...].shape[1] of THIS step: engines change it
                                                      (MJWarp: 0 while nothing touches; Genesis: per-step maximum) */
  int32_t ld_contact;                              /* elements between consecutive env rows of the three contact arrays */
  const uint64_t* contact_link_masks;              /* optional [N,2] {self-side, other-side} link bitmasks from
                                                      addk_contact_link_mask: used INSTEAD of the list when set */
} addk_sim_state;

/* Persistent per-env tensors owned by ADDObservation / ADDDone / Environment and mutated in place. */
typedef struct addk_env_buffers {
  float* time_buf;              /* [N]  env.time_buf */
  int64_t* motion_ids;          /* [N]  add_obs._motion_ids */
  float* motion_time_offsets;   /* [N]  add_obs._motion_time_offsets */
  float* ref_root_pos; float* ref_root_rot; float* ref_root_vel; float* ref_root_ang_vel; /* [N,3|4|3|3] */
  float* ref_dof_pos;  float* ref_dof_vel;  /* [N,D] */
  float* hist;                  /* [N, num_disc_steps, hist_stride] ring of sim states: pose half | vel half */
  int32_t hist_stride;
  float* obs_buf;               /* [N, obs_dim] */
  float* disc_obs;              /* [N, disc_obs_dim]   info["disc_obs"] */
  float* disc_obs_demo;         /* [N, disc_obs_dim]   info["disc_obs_demo"] */
  float* reward;                /* [N] */
  int32_t* done;                /* [N] add_done.done_buf */
  /* ReturnTracker (base_agent.py:564-621) */
  float* return_buf; int64_t* ep_len_buf; int64_t* eps_per_env;
  double* tracker_sums;         /* [2] sum of returns / of lengths over finished episodes */
  int64_t* tracker_count;       /* [1] finished episodes */
} addk_env_buffers;

/* Row `t` of the [T,N,...] experience buffers (ExperienceBuffer.record, experience_buffer.py:50-53). */
typedef struct addk_exp_row {
  float* next_obs; float* reward; int32_t* done; float* disc_obs; float* disc_obs_demo;
  int64_t* motion_ids; float* motion_times;
} addk_exp_row;

/* ----- motion table (csrc/motion.cu) --------------------------------------------------------- */
/* MotionLib._load_motion_pkl + _precompute_motion_steps for ONE clip (motion_lib.py:164-320).
 * frames [F,7+D] fp32 in file layout; writes rows [row0, row0+n_steps) of `table`. */
int addk_motion_table_build(void* stream, const float* frames, int num_frames, int num_dofs,
                            const int* col_of_dof, const float* dof_axis, float fps, float frame_dt,
                            int n_steps, double dt, float motion_len, int loop_wrap, float* jrot_work,
                            float* fvel_work, float* table, int row_stride, long long row0,
                            float* joint_rot_out, long long* frame_idx_out);
/* MotionLib.calc_motion_frame (motion_lib.py:61-88,118-150): interpolation at arbitrary (clip id, time) queries from
 * the 30 fps source frames of all clips concatenated -- frames [sumF, 7+D] (file layout), jrot [sumF, D, 4] and fvel
 * [sumF, 6+D] as stage A of addk_motion_table_build leaves them; frame_start [C] = cumulative 30 fps frame counts,
 * num_frames / lengths / loop_modes [C].  Outputs [n,3|4|3|3|D*4|D|D], any may be NULL. */
int addk_motion_frame(void* stream, const float* frames, const float* jrot, const float* fvel,
                      const long long* frame_start, const long long* num_frames, const float* lengths,
                      const int* loop_modes, int num_dofs, const float* dof_axis, const long long* ids,
                      const float* times, int n, float* root_pos, float* root_rot, float* root_vel,
                      float* root_ang_vel, float* joint_rot, float* dof_pos, float* dof_vel);
/* MotionLib.get_precomputed_motion_step (motion_lib.py:322-335); any output may be NULL. */
int addk_motion_gather(void* stream, const float* table, int row_stride, int num_dofs, long long s_total,
                       const long long* start_idx, float dt_inv, const long long* ids, const float* times,
                       int n, float* root_pos, float* root_rot, float* root_vel, float* root_ang_vel,
                       float* dof_pos, float* dof_vel, long long* idx_out);

/* ----- fused per-step work around the physics step (csrc/step.cu) ---------------------------- */
/* ADDAgent._step_env after scene.step() + _record_data_post_step + ReturnTracker.update
 * (add_agent.py:93-108,204-219; add_observation.py:163-207,296-306; add_reward.py:54-89;
 *  add_done.py:59-90; base_agent.py:596-621).  flags: bit0 advance time_buf by ctrl_dt,
 *  bit1 update ref motion + push history (update_motion), bit2 compute reward/done + tracker,
 *  bit3 only touch envs whose `env_mask` byte is non-zero.  `hist_head` is the ring slot written
 *  by this push (CircularBuffer._head before the push). `exp` may be NULL. */
int addk_env_step(void* stream, const addk_task* task_host, const addk_motion_lib* lib_host,
                  const addk_sim_state* sim_host, const addk_env_buffers* env_host,
                  const addk_exp_row* exp_host, const float* dof_err_w, const uint8_t* env_mask,
                  int num_envs, int hist_head, int flags);
/* ADDObservation.reset_idx for the envs with done != 0 (add_observation.py:308-344; env.py:157-163;
 * add_done.py:92-93): adopts (new_ids,new_times), zeroes time/done, refreshes ref_*, refills the
 * history from the table and emits the pose to write into the simulator.  reset_mask_out[N] = 1 for
 * the envs that were reset. `hist_head` is CircularBuffer._head. */
int addk_reset_done(void* stream, const addk_task* task_host, const addk_motion_lib* lib_host,
                    const addk_env_buffers* env_host, const long long* new_ids, const float* new_times,
                    int num_envs, int hist_head, int reset_all, int zero_time_done, float* qpos_out, float* qvel_out,
                    uint8_t* reset_mask_out);
/* ADDMotion.sample_time for every env (add_motion.py:53-61; motion_lib.py:35-39; sampler.py:57-92)
 * from three uniforms per env.  Candidates are only consumed where done != 0. */
int addk_sample_motion_time(void* stream, const float* motion_weights, int num_motions, const float* errors,
                            int num_segments, const float* seg_sizes, float dt, float min_start_time,
                            float temperature, int rand_reset, const int32_t* done, const float* uniforms,
                            int num_envs, unsigned int* temp_bits_work, long long* ids_out, float* times_out);
/* AdaptiveSegmentSampler.sample_start_frame for given clip ids and an explicit temperature (sampler.py:75-92);
 * uniforms [n,3], columns 1 and 2 are consumed. */
int addk_sample_start_time(void* stream, const float* errors, int num_segments, const float* seg_sizes, float dt,
                           float min_start_time, float temperature, const float* uniforms, int n,
                           const long long* clip_ids, float* times_out);
/* The arithmetic of AdaptiveSegmentSampler.sample_start_frame after its two draws (sampler.py:84-92):
 * t = seg * size[clip] + U * size[clip]; (t // dt) * dt; clamp(min = min_start_time) -- for callers that keep the
 * reference's torch.multinomial / torch.rand draws (bit-exact start times on identical draws). */
int addk_start_time_from_draws(void* stream, const float* seg_sizes, float dt, float min_start_time,
                               const long long* clip_ids, const long long* segments, const float* uniforms, int n,
                               float* times_out);
/* AdaptiveSegmentSampler.update_errors (sampler.py:20-55). sums/counts are [C*S] work buffers.
 * disc_obs_demo == NULL: disc_obs is a ready [n] vector of tracking errors. */
int addk_sampler_update_errors(void* stream, const long long* clip_ids, const float* timesteps,
                               const float* disc_obs, const float* disc_obs_demo, int disc_dim, int n,
                               const float* seg_sizes, int num_motions, int num_segments, double* sums_work,
                               int* counts_work, float* errors);

/* ----- engine-side hand-off (csrc/engine_side.cu) ------------------------- */
/* MJWarpEntity.get_contacts without the host round trip (engine/mjwarp_engine.py:896-986): flat contact arrays of the
 * backend (geom pairs [cap,2], world ids [cap], *nacon_dev valid entries -- read on the device) -> per world two uint64
 * link bitmasks {self-side bodies, other-side bodies} of the contacts between `self` and `other` (bit b = body id b;
 * ids >= 64 cannot be represented and are ignored).  link_masks_out [nworld, 2] is cleared by the call.  The step kernel
 * takes the result through addk_sim_state::contact_link_masks. */
int addk_contact_link_mask(void* stream, const int* geom_pairs, const int* world_ids, const int* nacon_dev,
                           int nacon_cap, const int* geom_bodyid, int ngeom, unsigned long long self_body_mask,
                           unsigned long long other_body_mask, int self_is_other, int exclude_self_contact, int nworld,
                           unsigned long long* link_masks_out);
/* The six state getters of robot.py:271-293 (MJWarp: mjwarp_engine.py:640-795) as ONE launch: qpos [nworld, ld_qpos] /
 * qvel [nworld, ld_qvel] -> packed rows [pos3 | quat4 wxyz | dof D | pad][vel3 | ang3 | dofvel D | pad] (the step
 * table's row format).  qpos_col [7+D] / qvel_col [6+D]: source column of every packed entry (-1 = 0.0). */
int addk_pack_state(void* stream, const float* qpos, int ld_qpos, const float* qvel, int ld_qvel, const int* qpos_col,
                    const int* qvel_col, int num_dofs, int nworld, float* rows_out, int row_stride);
/* PD-control prologue of MJWarpScene.step (mjwarp_engine.py:1565-1604): qfrc[w, dof_ids[j]] += clamp(kp_j (target_wj -
 * pos_wj) - kv_j vel_wj, +-max_torque) for local dofs j >= 6 with a non-zero gain; pos_col [n_local] = qpos column of
 * local dof j (-1: reads as 0, ball joints); max_torque <= 0: no clamp.  qfrc is zeroed by the caller. */
int addk_pd_control(void* stream, const float* qpos, int ld_qpos, const float* qvel, int ld_qvel, const float* target,
                    const float* kp, const float* kv, const int* pos_col, const int* dof_ids, int n_local,
                    float max_torque, int nworld, float* qfrc, int ld_qfrc);

/* ----- returns / advantages / statistics (csrc/gae.cu) --------------------------------------- */
/* compute_td_lambda_return + next_vals masking + adv (base_agent.py:624-647; ppo_agent.py:126-146). */
int addk_td_lambda(void* stream, const float* reward, const float* next_vals, const float* vals,
                   const int32_t* done, int T, int N, float discount, float td_lambda, float succ_val,
                   float fail_val, float* tar_val, float* adv);
/* The minibatch index window of ExperienceBuffer._sample_rand_idx (experience_buffer.py:90-113) in one launch:
 *   out[i] = perm[(head + i) mod perm_len] mod sample_count,  i < n.
 * The reference slices its device permutation, concatenates the tail with the head of the re-drawn permutation on
 * wrap-around (the tail is a VIEW, so it also shows the re-drawn values) and applies torch.remainder: three or four
 * library launches per optimizer step.  The caller re-draws `perm` in place BEFORE the call when head + n > perm_len. */
int addk_perm_window(void* stream, const long long* perm, long long perm_len, long long head, int n,
                     long long sample_count, long long* out_idx);
/* std_mean over rand_action_mask == 1 (unbiased) then clamp((adv-mean)/max(std,1e-5), +-clip)
 * (ppo_agent.py:147-153).  stats_out: [mean, std]. */
int addk_adv_normalize(void* stream, float* adv, const float* rand_action_mask, int n, float clip,
                       double* work3, float* stats_out);
/* _calc_disc_rewards tail + reward mix (amp_agent.py:194-206; add_agent.py:124):
 * r = w_task*task_r + w_disc * scale * -log(max(1 - sigmoid(logit), 1e-4)); stats_out [mean,std] of disc_r */
int addk_disc_reward(void* stream, const float* logits, float* reward_inout, int n, float scale, float w_task,
                     float w_disc, double* work3, float* stats_out);
/* Column sums over rows of X[n, dim] in fp64: mode 0 -> (sum, sum of squares), mode 1 -> sum |a - b|.
 * Replaces Normalizer.record / DiffNormalizer.record accumulated over one iteration
 * (normalizer.py:25-35; diff_normalizer.py:24-31). out is [2*dim] or [dim] doubles, pre-zeroed. */
int addk_column_stats(void* stream, const float* a, const float* b, long long n, int dim, int mode, double* out);
/* Normalizer.update / DiffNormalizer.update (normalizer.py:37-80; diff_normalizer.py:33-45). */
int addk_normalizer_update(void* stream, const double* sums, double new_count, int dim, int64_t* count,
                           float* mean, float* mean_sq, float* std, float min_var);
int addk_diff_normalizer_update(void* stream, const double* sum_abs, double new_count, int dim, int64_t* count,
                                float* mean_abs);

/* ----- dense layers on the CUDA cores / tensor cores (csrc/gemm.cu, csrc/gemm_tc.cu) ---------- */
/* C[M,N] = act( normA(A)[M,K] . op(B) + bias ) ; see csrc/gemm.cu for the flag bits. */
typedef struct addk_gemm_args {
  const float* A; int32_t lda;       /* A_T=0: [M,K] row-major ; A_T=1: [K,M] row-major */
  const float* B; int32_t ldb;       /* B_T=1: [N,K] row-major (nn.Linear weight) ; B_T=0: [K,N] */
  float* C; int32_t ldc;
  int32_t M, N, K;
  const float* bias;                 /* [N] or NULL */
  const float* a_mean; const float* a_std;  /* optional (A-mean)/std on load, per K column */
  const float* relu_mask_src; int32_t ld_mask;  /* optional: C *= (mask_src > 0), [M,N] */
  int32_t trans_a, trans_b, relu, split_k;  /* split_k > 1: slab z is written at C + z * slab_stride */
  int32_t accumulate;                /* C += result (single-slab only) */
  int64_t slab_stride;               /* floats between consecutive split-K slabs; 0 = M * ldc */
  /* precision "bf16" only: bf16 twins of the operands (same shapes / leading dimensions, in elements) and an
   * optional bf16 copy of the output.  NULL twins -> the call runs in tf32x3 on the fp32 operands. */
  const void* A16; const void* B16; void* C16;
  /* precision "f16x3" only (fp32-parity mode on the fp16 tensor pipe): A16 / B16 point at the fp16 "hi" plane of the
   * operand's twin (same shape / leading dimension as the fp32 operand, x*s rounded to fp16, s = the power of two that
   * puts max|x| into [2^14, 2^15)); the "lo" plane ((x*s - hi) * 2^11 rounded to fp16) starts a16_plane / b16_plane
   * ELEMENTS after it; a_amax / b_amax are device slots of two words {W, max}: max = the bit pattern of max|x|, W = a sticky scale
   * word that stays in force while max|x| * s(W) is inside [2^9, 2^15) (zero it if unused).  a16_ready / b16_ready = 0:
   * the call computes max|x| and writes both planes first (two extra launches per operand); != 0: the twin already
   * holds this operand (converted by an earlier call).  NULL twins -> the call runs in tf32x3. */
  int64_t a16_plane, b16_plane;
  uint32_t* a_amax; uint32_t* b_amax;
  int32_t a16_ready, b16_ready;      /* 0 convert (max pass + split pass) | 1 twin ready | 2 *amax valid: split pass only */
  uint32_t* c_amax;                  /* optional, single-slab calls: C's slot; max|C| as stored is atomicMax-ed into word [1], so
                                      * a later call that reads C can pass ready = 2; the caller zeroes word [1] first
                                      * (addk_f16x3_prep) */
  int64_t c16_plane;                 /* optional, with C16 + c_amax, N > 128: the epilogue also writes C's fp16 planes with the
                                      * scale of the slot's sticky word [0]; addk_f16x3_repair afterwards rewrites them in
                                      * the rare case that scale does not fit max|C|; then readers of C pass ready = 1 */
  /* 16-bit-only tensors (layers wider than 128 outputs, single slab: the persistent kernels).  no_f32 = 1: the fp32
   * output is NOT written -- C lives only as its 16-bit copy C16 (precision "bf16"); C may then be NULL.
   * relu_mask_src16: the ReLU-mask source as 16-bit values (bf16 copy / fp16 hi plane of a tensor that has no fp32
   * copy), pitch ld_mask ELEMENTS, used instead of relu_mask_src: element > 0 <=> sign clear and magnitude non-zero. */
  const void* relu_mask_src16;
  int32_t no_f32;
  /* ReLU masks as bit planes (precision "f16x3", persistent kernel, N a multiple of 128): a layer with relu = 1 and
   * relu_bits_out != NULL also leaves one bit per output element behind (set where the output is > 0; column
   * c = 64 g + 4 i + k, i < 16, k < 4, of a row lives in word [row * ld_bits + 2 g + (k >> 1)], bit 16 (k & 1) + i --
   * the order the store loop's warp ballots produce); a later layer masks its output with relu_bits_in instead of reading a whole
   * [M,N] mask tensor (relu_mask_src / relu_mask_src16 are then ignored): 1/32 .. 1/16 of the bytes, one 16-byte load
   * per accumulator row and tile instead of 32 latency-bound loads in the store loop.  ld_bits in 32-bit WORDS, a
   * multiple of 4; both pointers 16-byte aligned.  Only the kernels addk_gemm_is_persistent() reports honour them. */
  uint32_t* relu_bits_out;
  const uint32_t* relu_bits_in;
  int32_t ld_bits;
  /* Optional (persistent kernels, single slab, no accumulate, N and ldc multiples of 4): the epilogue also leaves the
   * column sums of every 32-row block of the output behind -- colsum_partials[(row / 32) * N + col], ceil(M / 32) rows of
   * N floats -- so that the bias gradient 1^T dY of a gradient tensor costs one small reduction over the partial rows
   * instead of a second pass over dY (csrc/mlp.cu: colsum_parts_kernel). */
  float* colsum_partials;
} addk_gemm_args;
/* 1 if addk_gemm would run this call on a persistent tensor-core kernel (the only ones that honour no_f32 /
 * relu_mask_src16), given 16-bit operands TMA can address */
int addk_gemm_is_persistent(const addk_gemm_args* args_host, int precision);
int addk_gemm(void* stream, const addk_gemm_args* args_host, int precision);
/* precision "f16x3": max|x| of the [rows, cols] fp32 tensor x (pitch ld) -> *amax_slot (bit pattern), then the two
 * fp16 planes hi16[0 .. rows*ld) and hi16[plane .. plane + rows*ld) described above.  Three stream-ordered operations. */
int addk_f16x3_convert(void* stream, const float* x, long long rows, int cols, int ld, void* hi16, long long plane,
                       uint32_t* amax_slot);
/* the split pass alone: *amax_slot already holds max|x| (left there by the kernel that produced x).  Optional:
 * colsum_partials [148 * 8, cols] floats -- the pass also leaves per-block column sums of x there (the bias gradient
 * 1^T dY for free) and *colsum_partial_rows (host int) says how many rows were written; 0 = shape not supported */
int addk_f16x3_split(void* stream, const float* x, long long rows, int cols, int ld, void* hi16, long long plane,
                     uint32_t* amax_slot, float* colsum_partials, int* colsum_partial_rows);
/* before a dense layer that leaves max|C| (c_amax) or the planes of its output (c16_plane) behind: clears the max word
 * and either folds the scale in force into the slot's sticky word (keep_sticky_word = 1, needed for c16_plane) or
 * clears it (0: the scale then follows max|x| alone and results do not depend on earlier calls); addk_f16x3_repair,
 * after the layer, rewrites the planes if the sticky scale turned out not to fit max|C| */
int addk_f16x3_prep(void* stream, uint32_t* slot, int keep_sticky_word);
/* the same for a table of n_total consecutive slots in one launch: slots [0, n_sticky) keep their history (sticky
 * word <- the scale in force, max word <- 0), the rest are cleared.  addk_update_minibatch and the inference entry
 * points issue it once per call instead of one single-thread launch in front of every layer. */
int addk_f16x3_prep_all(void* stream, uint32_t* slots, int n_sticky, int n_total);
int addk_f16x3_repair(void* stream, const float* x, long long rows, int cols, int ld, void* hi16, long long plane,
                      uint32_t* slot);

/* ----- PPO / ADD minibatch (csrc/mlp.cu) ------------------------------------------------------ */
struct addk_update_ctx;  /* opaque; plain host struct of device pointers, see csrc/mlp.cu */
int addk_update_ctx_size(void);
/* fills in the context from a flat pointer table; layout documented in add_gym_b200/_lib.py */
int addk_update_ctx_init(void* ctx_host, const void* const* ptrs, int n_ptrs, const int64_t* ints, int n_ints,
                         const double* floats, int n_floats);
/* one optimizer step: ExperienceBuffer.sample gather + AMPAgent._compute_loss + backward + AdamW
 * (experience_buffer.py:74-113; amp_agent.py:98-114; ppo_agent.py:194-275; add_agent.py:141-202;
 *  mp_optimizer.py:14-23).  idx: the int64 minibatch permutation slice. */
int addk_update_minibatch(void* stream, void* ctx_host, const long long* idx, int step_index, int do_optim);
/* The derived copies of the parameters that the inference entry points (addk_actor_step / addk_critic_eval /
 * addk_disc_eval) otherwise rebuild on every call: the 16-bit twin of the flat parameter vector and the first-layer
 * weights with padded rows.  The reference has no counterpart (its nn.Linear weights are read directly,
 * learning/ppo_model.py:13-21); here 32 env steps of one rollout share one conversion: call this once after the
 * last parameter change and pass a context whose params16_current field is 1 to those entry points. */
int addk_params_refresh(void* stream, void* ctx_host);

/* actor inference for one env step (ppo_agent.py:72-104) */
int addk_actor_step(void* stream, void* ctx_host, const float* obs, const float* noise, const float* exp_mask,
                    int n, float* action, float* a_logp, float* obs_rec, float* action_rec, float* logp_rec,
                    float* mask_rec);
/* critic over rows (ppo_agent.py:120-144) and discriminator logits (amp_agent.py:194-200) */
int addk_critic_eval(void* stream, void* ctx_host, const float* obs, long long n, float* vals);
int addk_disc_eval(void* stream, void* ctx_host, const float* disc_obs, const float* disc_obs_demo, long long n,
                   float* logits);
/* AdamW on a flat fp32 parameter vector (torch.optim.AdamW, mp_optimizer.py:38; lr,betas,eps,wd). */
int addk_adamw(void* stream, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
               int step, double lr, double beta1, double beta2, double eps, double weight_decay,
               double grad_scale);

/* ----- the exchange step over NVLink peer memory (csrc/p2p.cu) ---------------------------------- */
/* Device memory the peers can map (cudaMalloc, zeroed) and its 64-byte cudaIpc handle; addk_p2p_open maps a peer's. */
int addk_p2p_alloc(long long bytes, void** ptr_out);
int addk_p2p_free(void* ptr);
int addk_p2p_export(void* ptr, unsigned char* handle64_host);
int addk_p2p_open(const unsigned char* handle64_host, void** ptr_out);
/* Gradient reduce-scatter + AdamW on this rank's shard + all-gather of the updated parameters in ONE kernel (replaces
 * dist.all_reduce(flat_grad) + addk_adamw; the all-reduce the reference's DDP would issue, base_agent.py:47-57).
 * grad / param / flag pointers of all `world` ranks in rank order (own entries included; flags: [2][8] uint32, zeroed).
 * `n` floats (vectors padded to a multiple of 4); `step` = optimizer step number, identical on all ranks (it is the
 * epoch of the flags); only this rank's shard of exp_avg / exp_avg_sq is updated.  ticket: one zeroed uint32. */
int addk_p2p_adamw(void* stream, int rank, int world, float* const* grad_ptrs_host, float* const* param_ptrs_host,
                   unsigned int* const* flag_ptrs_host, float* exp_avg, float* exp_avg_sq, long long n, int step,
                   double lr, double beta1, double beta2, double eps, double weight_decay, unsigned int* ticket,
                   int max_blocks);

/* MPOptimizer._clip_grads = torch.nn.utils.clip_grad_norm_ (mp_optimizer.py:19-20,46-47) on the flat gradient:
 * coef = min(1, max_norm / (pre_scale * ||g||_2 + 1e-6)); g *= coef.  pre_scale = 1/world while g holds the cross-rank
 * sum.  sumsq_work: one double; *coef_out receives the coefficient. */
int addk_clip_grad_norm(void* stream, float* grads, long long n, double max_norm, double pre_scale,
                        double* sumsq_work, float* coef_out);

#ifdef __cplusplus
}
#endif
#endif /* ADDK_H_ */
