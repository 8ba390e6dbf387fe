"""CPU restatement of the engine-side hand-off the reference does in Python.  TEST INFRASTRUCTURE -- NOT THE PRODUCT.

Each function follows the cited reference lines (paths relative to /root/reference/add_gym/) on plain torch / numpy data;
tests/test_gpu_engine_side.py compares csrc/engine_side.cu with them on synthetic inputs (MuJoCo is not installable here).
Pinning: the reference functions need a compiled MuJoCo model to run, so they cannot be executed in this container --
parity for these three is "restated, unpinned"; the restatement is a line-by-line port of loops without arithmetic
subtleties (index selection, one multiply-subtract-clamp)."""
import numpy as np
import torch


def get_contacts(geom_pairs, world_ids, nacon, geom_bodyid, self_bodies, other_bodies, self_is_other, exclude_self_contact,
                 nworld):
    """MJWarpEntity.get_contacts (engine/mjwarp_engine.py:896-986): -> {link_a, link_b, valid_mask} padded to the
    per-call maximum, [nworld, 0] when there is no contact."""
    if nacon <= 0:
        e = torch.empty((nworld, 0), dtype=torch.long)
        return {"link_a": e, "link_b": e.clone(), "valid_mask": torch.empty((nworld, 0), dtype=torch.bool)}
    gp = np.asarray(geom_pairs[:nacon], dtype=np.int32)
    wi = np.asarray(world_ids[:nacon], dtype=np.int32)
    gb = np.asarray(geom_bodyid, dtype=np.int32)
    self_set, other_set = set(int(b) for b in self_bodies), set(int(b) for b in other_bodies)
    pa = [[] for _ in range(nworld)]
    pb = [[] for _ in range(nworld)]
    for i in range(nacon):                                                  # :929-951
        w = int(wi[i])
        g0, g1 = int(gp[i, 0]), int(gp[i, 1])
        if g0 < 0 or g1 < 0:
            continue
        b0, b1 = int(gb[g0]), int(gb[g1])
        if exclude_self_contact and self_is_other:
            if b0 in self_set and b1 in self_set:
                continue
        a_in_self, b_in_self = b0 in self_set, b1 in self_set
        a_in_other, b_in_other = b0 in other_set, b1 in other_set
        if a_in_self and b_in_other:
            pa[w].append(b0); pb[w].append(b1)
        elif b_in_self and a_in_other:
            pa[w].append(b1); pb[w].append(b0)
    max_len = max((len(x) for x in pa), default=0)                          # :953-984
    link_a = torch.full((nworld, max_len), -1, dtype=torch.long)
    link_b = torch.full((nworld, max_len), -1, dtype=torch.long)
    valid = torch.zeros((nworld, max_len), dtype=torch.bool)
    for w in range(nworld):
        n = len(pa[w])
        if n:
            link_a[w, :n] = torch.tensor(pa[w]); link_b[w, :n] = torch.tensor(pb[w]); valid[w, :n] = True
    return {"link_a": link_a, "link_b": link_b, "valid_mask": valid}


def contact_bool(contacts, link_ids):
    """Manipulator.get_ground_contact_forces_v2 (robot.py:221-231)."""
    ids = torch.as_tensor(link_ids, dtype=torch.long)
    a = torch.isin(contacts["link_a"], ids) & contacts["valid_mask"]
    b = torch.isin(contacts["link_b"], ids) & contacts["valid_mask"]
    return a.any(dim=1) | b.any(dim=1)


def link_masks(contacts, nworld):
    """The per-world {link_a bodies, link_b bodies} bitmasks addk_contact_link_mask produces, from the padded lists."""
    out = np.zeros((nworld, 2), dtype=np.uint64)
    la, lb, va = (contacts[k].numpy() for k in ("link_a", "link_b", "valid_mask"))
    for w in range(nworld):
        for j in range(la.shape[1]):
            if va[w, j]:
                if 0 <= la[w, j] < 64:
                    out[w, 0] |= np.uint64(1) << np.uint64(la[w, j])
                if 0 <= lb[w, j] < 64:
                    out[w, 1] |= np.uint64(1) << np.uint64(lb[w, j])
    return out


def get_dofs_position(qpos, segments, n_dofs):
    """MJWarpEntity.get_dofs_position (engine/mjwarp_engine.py:712-731).  segments: (kind, dof_local_adr, qpos_adr)."""
    out = torch.zeros((qpos.shape[0], n_dofs), dtype=qpos.dtype)
    for kind, dadr, qadr in segments:
        if kind == "free":
            out[:, dadr:dadr + 3] = qpos[:, qadr:qadr + 3]
        elif kind == "ball":
            continue
        else:
            out[:, dadr] = qpos[:, qadr]
    return out


def packed_state(qpos, qvel, segments, dof_ids, n_dofs, free_qpos_adr, row_stride):
    """robot.py:271-293 over the MJWarp getters (mjwarp_engine.py:640-795): base pos / quat (wxyz) / lin vel / ang vel /
    dof_pos[:, 6:] / dof_vel[:, 6:] laid out as the packed row [pos3 quat4 dof D pad | vel3 ang3 dofvel D pad]."""
    D = n_dofs - 6
    half = (7 + D + 3) & ~3
    out = torch.zeros((qpos.shape[0], row_stride), dtype=torch.float32)
    dpos = get_dofs_position(qpos, segments, n_dofs)
    dvel = qvel[:, torch.as_tensor(dof_ids, dtype=torch.long)]              # get_dofs_velocity :793-795
    out[:, 0:3] = qpos[:, free_qpos_adr:free_qpos_adr + 3]                  # get_pos
    out[:, 3:7] = qpos[:, free_qpos_adr + 3:free_qpos_adr + 7]              # get_quat (MuJoCo stores wxyz)
    out[:, 7:7 + D] = dpos[:, 6:]
    out[:, half:half + 3] = dvel[:, 0:3]                                    # get_vel :680-694
    out[:, half + 3:half + 6] = dvel[:, 3:6]                                # get_ang :696-709
    out[:, half + 6:half + 6 + D] = dvel[:, 6:]
    return out


def pd_control(qpos, qvel, target, kp, kv, segments, dof_ids, n_dofs, max_torque, qfrc):
    """PD prologue of MJWarpScene.step (engine/mjwarp_engine.py:1565-1604), one substep; qfrc is zeroed first."""
    qfrc = torch.zeros_like(qfrc)
    if float(kp.abs().max().item()) == 0.0 and float(kv.abs().max().item()) == 0.0:
        return qfrc
    pos = get_dofs_position(qpos, segments, n_dofs)
    ids = torch.as_tensor(dof_ids, dtype=torch.long)
    vel = qvel[:, ids]
    tau = kp.unsqueeze(0) * (target - pos) - kv.unsqueeze(0) * vel
    if max_torque is not None and max_torque > 0:
        tau = torch.clamp(tau, -max_torque, max_torque)
    mask = (kp != 0) | (kv != 0)
    if mask.numel() >= 6:
        mask[:6] = False
    if mask.any():
        qfrc[:, ids[mask]] += tau[:, mask]
    return qfrc
