// Field table of addk_update_ctx (X-macros).  The C++ side expands it into the struct and into the
// loader behind addk_update_ctx_init(); add_gym_b200/_lib.py parses this very file to learn the order,
// so the two sides cannot drift apart.
//
//   ADDK_PTR(name)   device pointer       ADDK_INT(name)   int64      ADDK_F64(name)   double
// clang-format off
#ifndef ADDK_PTR
#define ADDK_PTR(n)
#endif
#ifndef ADDK_INT
#define ADDK_INT(n)
#endif
#ifndef ADDK_F64
#define ADDK_F64(n)
#endif

// ---- flat parameter vector and optimizer state (22 trainable tensors, reference order) ----------
ADDK_PTR(params)        // [P] fp32
ADDK_PTR(grads)         // [P] fp32, reduced gradient of the last minibatch
ADDK_PTR(slabs)         // [2*split_k, P] split-K partial weight gradients
ADDK_PTR(exp_avg)       // [P]
ADDK_PTR(exp_avg_sq)    // [P]
// ---- normalizers ----------------------------------------------------------------------------------
ADDK_PTR(obs_mean)      // [obs_dim]
ADDK_PTR(obs_std)
ADDK_PTR(a_mean)        // [act_dim]
ADDK_PTR(a_std)
ADDK_PTR(disc_mean_abs) // [disc_dim]
ADDK_PTR(logstd)        // [act_dim] _action_dist._logstd_net (frozen)
// ---- flat [T*N, ...] experience buffers -------------------------------------------------------------
ADDK_PTR(buf_obs)
ADDK_PTR(buf_action)
ADDK_PTR(buf_a_logp)
ADDK_PTR(buf_adv)
ADDK_PTR(buf_tar_val)
ADDK_PTR(buf_mask)
ADDK_PTR(buf_disc_obs)
ADDK_PTR(buf_disc_demo)
// ---- minibatch workspace (R = mb_rows + 1 rows; the extra row is the discriminator's "zero diff" sample)
ADDK_PTR(xn)            // [R, obs_ld]      normalised obs (columns >= obs_dim are zero)
ADDK_PTR(an)            // [R, act_ld]      normalised action
ADDK_PTR(old_logp)      // [R]
ADDK_PTR(adv)           // [R]
ADDK_PTR(tar)           // [R]
ADDK_PTR(mask)          // [R]
ADDK_PTR(dn)            // [R, disc_ld]     normalised (demo - agent) difference, row mb_rows = 0
ADDK_PTR(h1)            // [R, 1024]
ADDK_PTR(h2)            // [R, 1024]
ADDK_PTR(h3)            // [R, 512]
ADDK_PTR(g1)            // [R, 1024]
ADDK_PTR(g2)            // [R, 1024]
ADDK_PTR(g3)            // [R, 512]
ADDK_PTR(u1)            // [R, 1024]
ADDK_PTR(u2)            // [R, 512]
ADDK_PTR(gx)            // [R, disc_ld]
ADDK_PTR(dg)            // [R, disc_ld]
ADDK_PTR(mean)          // [R, act_ld]
ADDK_PTR(dmean)         // [R, act_ld]
ADDK_PTR(pred)          // [R]
ADDK_PTR(dpred)         // [R]
ADDK_PTR(ones)          // [R] all 1.0f
ADDK_PTR(stats)         // [32] double accumulators of the current minibatch
ADDK_PTR(info)          // [max_steps, 16] float diagnostics, one row per optimizer step
ADDK_PTR(cnt)           // [1] int: rows of the minibatch with rand_action_mask == 1
ADDK_PTR(colsum_work)   // [128*1024 + 64] floats: column-sum partials + ticket counters (zero-initialised)
ADDK_PTR(wd0_pad)       // [hid_d1, disc_ld] discriminator first-layer weight with rows padded to 16 bytes
ADDK_PTR(wa0_pad)       // [hid_a1, obs_ld] actor / critic first-layer weights with rows padded to obs_ld (tensor-core modes)
ADDK_PTR(wc0_pad)
// ---- precision "bf16": every fp32 workspace tensor above is carved out of ONE arena, so the bf16 twin of any
//      operand is arena16 + (ptr - arena); params16 is the bf16 shadow of the flat parameter vector
ADDK_PTR(arena)         // fp32 arena base (may be NULL when precision != bf16)
ADDK_PTR(arena16)       // bf16 arena, same element offsets
ADDK_PTR(params16)      // [P] bf16
// ---- precision "f16x3": arena16 / params16 hold TWO fp16 planes each (hi, then lo arena_elems / num_params elements
//      later) and amax_slots the max|x| words of the twins (slot 0 = the parameter vector); see csrc/gemm_tc.cu
ADDK_PTR(amax_slots)    // [2 * (1 + 4 * 128)] uint32: {sticky scale word, max|x|} per twin (NULL unless precision == f16x3)
ADDK_PTR(arena_bits)    // [arena_elems / 32] uint32: ReLU bit planes, the mask of the arena tensor at element offset o starts at word o / 32
                        //   (arena pieces start at multiples of 128 elements; NULL: masks are read from the tensors themselves)
// ---- second and third workspace sets: the critic and the discriminator chains of one optimizer step run on their
//      own streams next to the actor's (n_streams == 3), so the tail wave of one chain's dense layer overlaps the
//      next chain's tiles; all NULL / n_streams == 1 = the three chains run back to back on the caller's stream
ADDK_PTR(c_h1) ADDK_PTR(c_h2) ADDK_PTR(c_h3) ADDK_PTR(c_g1) ADDK_PTR(c_g2) ADDK_PTR(c_g3)
ADDK_PTR(d_e1)          // [R, hid_d1]
ADDK_PTR(d_e2)          // [R, hid_d2]
ADDK_PTR(d_dh2)         // [R, hid_d2]
ADDK_PTR(d_dv1)         // [R, hid_d1]
ADDK_PTR(d_du2)         // [R, hid_d2]
ADDK_PTR(d_pred)        // [R]
ADDK_PTR(d_dpred)       // [R]
ADDK_PTR(colsum_work_c) // like colsum_work, for the critic's stream
ADDK_PTR(colsum_work_d) // like colsum_work, for the discriminator's stream
// ---- precision "f16x3": per-block column sums the split pass of a gradient tensor leaves behind (bias gradients), one
//      buffer per chain: [148 * 8, 1024] floats each (NULL: the separate column-sum kernel runs)
ADDK_PTR(colpart_a) ADDK_PTR(colpart_c) ADDK_PTR(colpart_d)

ADDK_INT(obs_dim)
ADDK_INT(act_dim)
ADDK_INT(disc_dim)
ADDK_INT(obs_ld)        // obs_dim rounded up to a multiple of 16: row pitch of xn and of the padded first-layer weights
ADDK_INT(act_ld)        // act_dim rounded up to a multiple of 8
ADDK_INT(disc_ld)       // disc_dim rounded up to a multiple of 16 (32-byte row pitch in the 16-bit twins)
ADDK_INT(mb_rows)       // minibatch rows M
ADDK_INT(num_params)    // P (including alignment padding)
ADDK_INT(split_k)
ADDK_INT(arena_elems)   // elements in the arena
ADDK_INT(precision)     // 0 fp32 | 1 tf32x3 | 2 tf32 | 3 bf16 | 4 f16x3
ADDK_INT(n_streams)     // 1 | 3
ADDK_INT(params16_current) // 1: the caller guarantees that nothing changed the parameters since addk_params_refresh() ran on this
                        //    context's buffers -- the inference entry points then skip the per-call conversion of the flat
                        //    parameter vector (max pass + split pass / bf16 copy, 17 MB) and the re-padding of the first-layer
                        //    weights (32 env steps of one rollout share one set of weights).  addk_update_minibatch refuses it.
ADDK_INT(hid_a1)        // actor/critic hidden sizes (1024, 1024, 512)
ADDK_INT(hid_a2)
ADDK_INT(hid_a3)
ADDK_INT(hid_d1)        // discriminator hidden sizes (1024, 512)
ADDK_INT(hid_d2)
// offsets (in floats) of the 22 tensors inside the flat vector
ADDK_INT(o_a_w0) ADDK_INT(o_a_b0) ADDK_INT(o_a_w1) ADDK_INT(o_a_b1) ADDK_INT(o_a_w2) ADDK_INT(o_a_b2)
ADDK_INT(o_a_wm) ADDK_INT(o_a_bm)
ADDK_INT(o_c_w0) ADDK_INT(o_c_b0) ADDK_INT(o_c_w1) ADDK_INT(o_c_b1) ADDK_INT(o_c_w2) ADDK_INT(o_c_b2)
ADDK_INT(o_c_wo) ADDK_INT(o_c_bo)
ADDK_INT(o_d_w0) ADDK_INT(o_d_b0) ADDK_INT(o_d_w1) ADDK_INT(o_d_b1) ADDK_INT(o_d_wl) ADDK_INT(o_d_bl)

ADDK_F64(ppo_clip_ratio)
ADDK_F64(action_bound_weight)
ADDK_F64(critic_loss_weight)
ADDK_F64(disc_loss_weight)
ADDK_F64(disc_logit_reg)
ADDK_F64(disc_grad_penalty)
ADDK_F64(disc_weight_decay)
ADDK_F64(lr)
ADDK_F64(beta1)
ADDK_F64(beta2)
ADDK_F64(adam_eps)
ADDK_F64(weight_decay)
ADDK_F64(grad_scale)    // 1/world_size when gradients were summed across ranks, else 1
ADDK_F64(grad_clip)     // optimizer.grad_clip (mp_optimizer.py:10): global-norm clip before the step; 0 = off (the reference default, Q4)

#undef ADDK_PTR
#undef ADDK_INT
#undef ADDK_F64
// clang-format on
