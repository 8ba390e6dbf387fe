"""GPU: the engine-side hand-off kernels (csrc/engine_side.cu) against the restated reference loops
(oracle/engine_side_oracle.py) on synthetic backend arrays -- no MuJoCo needed."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(D=29, extra_qpos=5):
    """A MuJoCo-like layout: a few unrelated qpos/qvel entries first (another entity), then a free joint, then D scalar
    joints in MODEL order; the entity's local dof order (BFS) is a permutation of them, as in mjwarp_engine.py:1500-1540."""
    rng = np.random.default_rng(0)
    free_qpos, free_dof = extra_qpos, extra_qpos - 1
    perm = rng.permutation(D)
    segments = [("free", 0, free_qpos)]
    dof_ids = [free_dof + i for i in range(6)]
    for local, j in enumerate(perm):
        segments.append(("hinge", 6 + local, free_qpos + 7 + int(j)))
        dof_ids.append(free_dof + 6 + int(j))
    nq, nv = free_qpos + 7 + D + 2, free_dof + 6 + D + 2
    return segments, dof_ids, 6 + D, free_qpos, nq, nv


def test_contact_link_masks_match_the_reference_loop_and_the_step_kernel_accepts_them():
    from add_gym_b200 import _lib
    from oracle import engine_side_oracle as ref
    L = _lib.lib()
    rng = np.random.default_rng(1)
    nworld, ngeom, cap = 300, 70, 4096
    geom_bodyid = rng.integers(0, 34, size=ngeom).astype(np.int32)
    geom_bodyid[0] = 0                                               # the plane's geom: world body 0
    self_bodies, other_bodies = list(range(1, 31)), [0]
    for nacon in (0, 1, 777, cap):
        gp = rng.integers(-1, ngeom, size=(cap, 2)).astype(np.int32)
        gp[rng.random(cap) < 0.5, 1] = 0                             # half of the contacts are against the plane
        wi = rng.integers(0, nworld, size=cap).astype(np.int32)
        for self_is_other, other, excl in ((0, other_bodies, 1), (1, self_bodies, 1), (1, self_bodies, 0)):
            contacts = ref.get_contacts(gp, wi, nacon, geom_bodyid, self_bodies, other, bool(self_is_other), bool(excl), nworld)
            want = ref.link_masks(contacts, nworld)
            out = torch.full((nworld, 2), -1, dtype=torch.int64, device="cuda")
            sm = sum(1 << b for b in self_bodies)
            om = sum(1 << b for b in other)
            # (device copies are held in variables: a temporary would go back to the caching allocator before the launch runs)
            d_gp, d_wi, d_gb = torch.from_numpy(gp).cuda(), torch.from_numpy(wi).cuda(), torch.from_numpy(geom_bodyid).cuda()
            d_n = torch.tensor([nacon], dtype=torch.int32, device="cuda")
            rc = L.addk_contact_link_mask(_lib.stream(), _lib.ptr(d_gp), _lib.ptr(d_wi), _lib.ptr(d_n), C.c_int(cap),
                                          _lib.ptr(d_gb), C.c_int(ngeom), C.c_ulonglong(sm),
                                          C.c_ulonglong(om), C.c_int(self_is_other), C.c_int(excl), C.c_int(nworld), _lib.ptr(out))
            _lib.check(rc, "addk_contact_link_mask")
            got = out.cpu().numpy().view(np.uint64)
            assert np.array_equal(got, want), (nacon, self_is_other, excl)
            # the consumer's question (robot.py:221-231) answered from the masks == answered from the padded lists
            noncontact = [b for b in self_bodies if b not in (6, 12)]
            nm = np.uint64(sum(1 << b for b in noncontact))
            from_masks = ((got[:, 0] | got[:, 1]) & nm) != 0
            assert np.array_equal(from_masks, ref.contact_bool(contacts, noncontact).numpy())


def test_step_kernel_takes_link_masks_instead_of_the_contact_list():
    """Same synthetic contacts handed to the fused step kernel as the padded list and as link bitmasks: identical done flags."""
    from add_gym_b200 import config as b200_config
    from add_gym_b200.add_agent import ADDAgent

    def run(use_masks):
        cfg = b200_config.default_config(num_envs=512)
        cfg["engine"].update(seed=7, noise_device="device", fall_prob=0.05)
        cfg["agent"]["cuda_graphs"] = False
        torch.manual_seed(0)
        torch.cuda.manual_seed(0)
        a = ADDAgent(cfg, device="cuda:0")
        ent = a._env.robot.entity
        if use_masks:
            def link_masks(with_entity=None, exclude_self_contact=True):
                c = ent.get_contacts(with_entity=with_entity, exclude_self_contact=exclude_self_contact)
                la, lb, va = c["link_a"].long(), c["link_b"].long(), c["valid_mask"].bool()
                one = torch.ones_like(la)
                ba = torch.where(va, one << la.clamp(0, 62), torch.zeros_like(la))
                bb = torch.where(va, one << lb.clamp(0, 62), torch.zeros_like(lb))
                ma, mb = ba[:, 0], bb[:, 0]
                for j in range(1, ba.shape[1]):                # bitwise OR over the slots (several slots may name one link)
                    ma, mb = ma | ba[:, j], mb | bb[:, j]
                return torch.stack([ma, mb], dim=1).contiguous()
            ent.get_contact_link_masks = link_masks
        a._curr_obs, a._curr_info = a._reset_envs()
        a._exp_buffer.clear()
        a._rollout_train(8)
        return a._exp_buffer.get_data("done")[:8].clone()

    d0, d1 = run(False), run(True)
    assert int((d0 == 1).sum()) > 20
    assert torch.equal(d0, d1), (int((d0 != d1).sum()), (d0 != d1).nonzero()[:5].tolist(), int((d0 == 1).sum()), int((d1 == 1).sum()))


def test_pack_state_and_pd_control_match_the_reference_getters():
    from add_gym_b200 import _lib
    from oracle import engine_side_oracle as ref
    L = _lib.lib()
    D = 29
    segments, dof_ids, n_dofs, free_qpos, nq, nv = _model(D)
    g = torch.Generator().manual_seed(2)
    nworld = 1000
    qpos = torch.randn(nworld, nq, generator=g)
    qvel = torch.randn(nworld, nv, generator=g)
    # ---- pack: the six getters as one launch
    half = (7 + D + 3) & ~3
    want = ref.packed_state(qpos, qvel, segments, dof_ids, n_dofs, free_qpos, 2 * half)
    hinge = {dadr: qadr for kind, dadr, qadr in segments if kind != "free"}
    qpos_col = [free_qpos + i for i in range(7)] + [hinge[6 + j] for j in range(D)]
    qvel_col = list(dof_ids)
    rows = torch.full((nworld, 2 * half), float("nan"), device="cuda")
    d_qpos, d_qvel = qpos.cuda(), qvel.cuda()
    d_pc, d_vc = torch.tensor(qpos_col, dtype=torch.int32, device="cuda"), torch.tensor(qvel_col, dtype=torch.int32, device="cuda")
    rc = L.addk_pack_state(_lib.stream(), _lib.ptr(d_qpos), C.c_int(nq), _lib.ptr(d_qvel), C.c_int(nv), _lib.ptr(d_pc),
                           _lib.ptr(d_vc), C.c_int(D), C.c_int(nworld), _lib.ptr(rows), C.c_int(2 * half))
    _lib.check(rc, "addk_pack_state")
    assert torch.equal(rows.cpu(), want), "pure copies: bit-exact"
    # ---- PD prologue
    for max_torque, zero_gains in ((0.0, False), (35.0, False), (35.0, True)):
        kp = torch.rand(n_dofs, generator=g) * 100.0
        kv = torch.rand(n_dofs, generator=g) * 5.0
        kp[9] = 0.0; kv[9] = 0.0                                       # a dof without gains is skipped
        kp[11] = 0.0                                                   # damping only
        if zero_gains:
            kp.zero_(); kv.zero_()
        target = torch.randn(nworld, n_dofs, generator=g)
        want = ref.pd_control(qpos, qvel, target, kp, kv, segments, dof_ids, n_dofs, max_torque, torch.zeros(nworld, nv))
        pos_col = [free_qpos, free_qpos + 1, free_qpos + 2, -1, -1, -1] + [hinge[6 + j] for j in range(D)]
        qfrc = torch.zeros(nworld, nv, device="cuda")
        d_t, d_kp, d_kv = target.cuda(), kp.cuda(), kv.cuda()
        d_pc2, d_ids = torch.tensor(pos_col, dtype=torch.int32, device="cuda"), torch.tensor(dof_ids, dtype=torch.int32, device="cuda")
        rc = L.addk_pd_control(_lib.stream(), _lib.ptr(d_qpos), C.c_int(nq), _lib.ptr(d_qvel), C.c_int(nv), _lib.ptr(d_t),
                               _lib.ptr(d_kp), _lib.ptr(d_kv), _lib.ptr(d_pc2), _lib.ptr(d_ids), C.c_int(n_dofs),
                               C.c_float(max_torque), C.c_int(nworld), _lib.ptr(qfrc), C.c_int(nv))
        _lib.check(rc, "addk_pd_control")
        assert torch.equal(qfrc.cpu(), want), "kp (t - p) - kv v with separately rounded products: bit-exact"
