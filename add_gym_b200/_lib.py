"""ctypes binding of libaddk.so (the C-ABI declared in include/addk.h).

There is no fallback: if the shared library is missing, importing a kernel entry point raises, and
every wrapper refuses tensors that are not on a CUDA device.  PyTorch is used for device memory,
streams and torch.distributed only; all hot-path arithmetic happens inside the library.
"""
import ctypes as C
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# ADDK_LIB selects another build of the same C-ABI (e.g. libaddk_legacy.so from `make LEGACY=1`, which adds the superseded
# "tf32" / "tf32x3" precision modes); it is read once, at import
LIB_PATH = os.environ.get("ADDK_LIB") or os.path.join(_HERE, "libaddk.so")
CTX_FIELDS_PATH = os.path.join(_HERE, "csrc", "ctx_fields.h")

ADDK_MAX_TAR_STEPS = 16
ADDK_MAX_DISC_STEPS = 8

F_ADVANCE, F_UPDATE_MOTION, F_REWARD_DONE, F_MASKED = 1, 2, 4, 8
PRECISIONS = {"fp32": 0, "tf32x3": 1, "tf32": 2, "bf16": 3, "f16x3": 4}


class AddkTask(C.Structure):
    _fields_ = [
        ("num_dofs", C.c_int32), ("num_tar_steps", C.c_int32), ("num_disc_steps", C.c_int32),
        ("global_obs", C.c_int32), ("root_height_obs", C.c_int32), ("enable_vel_obs", C.c_int32),
        ("enable_phase_obs", C.c_int32), ("enable_tar_obs", C.c_int32), ("num_phase_encoding", C.c_int32),
        ("obs_dim", C.c_int32), ("disc_obs_dim", C.c_int32), ("track_root", C.c_int32), ("track_root_h", C.c_int32),
        ("enable_early_termination", C.c_int32), ("pose_termination", C.c_int32), ("contact_slots", C.c_int32),
        ("tar_offsets", C.c_float * ADDK_MAX_TAR_STEPS), ("disc_offsets", C.c_float * ADDK_MAX_DISC_STEPS),
        ("ctrl_dt", C.c_float), ("dt_inv", C.c_float),
        ("pose_w", C.c_float), ("vel_w", C.c_float), ("root_pose_w", C.c_float), ("root_vel_w", C.c_float),
        ("pose_scale", C.c_float), ("vel_scale", C.c_float), ("root_pose_scale", C.c_float),
        ("root_vel_scale", C.c_float), ("ep_len", C.c_float), ("pose_termination_dist", C.c_float),
        ("noncontact_link_mask", C.c_uint64),
    ]


class AddkMotionLib(C.Structure):
    _fields_ = [("table", C.c_void_p), ("row_stride", C.c_int32), ("num_motions", C.c_int32), ("s_total", C.c_int64),
                ("start_idx", C.c_void_p), ("lengths", C.c_void_p), ("loop_modes", C.c_void_p)]


class AddkSimState(C.Structure):
    _fields_ = [("root_pos", C.c_void_p), ("ld_root_pos", C.c_int32), ("root_rot", C.c_void_p), ("ld_root_rot", C.c_int32),
                ("root_vel", C.c_void_p), ("ld_root_vel", C.c_int32), ("root_ang", C.c_void_p), ("ld_root_ang", C.c_int32),
                ("dof_pos", C.c_void_p), ("ld_dof_pos", C.c_int32), ("dof_vel", C.c_void_p), ("ld_dof_vel", C.c_int32),
                ("link_a", C.c_void_p), ("link_b", C.c_void_p), ("valid", C.c_void_p),
                ("contact_slots", C.c_int32), ("ld_contact", C.c_int32), ("contact_link_masks", C.c_void_p)]


class AddkEnvBuffers(C.Structure):
    _fields_ = [("time_buf", C.c_void_p), ("motion_ids", C.c_void_p), ("motion_time_offsets", C.c_void_p),
                ("ref_root_pos", C.c_void_p), ("ref_root_rot", C.c_void_p), ("ref_root_vel", C.c_void_p),
                ("ref_root_ang_vel", C.c_void_p), ("ref_dof_pos", C.c_void_p), ("ref_dof_vel", C.c_void_p),
                ("hist", C.c_void_p), ("hist_stride", C.c_int32), ("obs_buf", C.c_void_p), ("disc_obs", C.c_void_p),
                ("disc_obs_demo", C.c_void_p), ("reward", C.c_void_p), ("done", C.c_void_p),
                ("return_buf", C.c_void_p), ("ep_len_buf", C.c_void_p), ("eps_per_env", C.c_void_p),
                ("tracker_sums", C.c_void_p), ("tracker_count", C.c_void_p)]


class AddkExpRow(C.Structure):
    _fields_ = [("next_obs", C.c_void_p), ("reward", C.c_void_p), ("done", C.c_void_p), ("disc_obs", C.c_void_p),
                ("disc_obs_demo", C.c_void_p), ("motion_ids", C.c_void_p), ("motion_times", C.c_void_p)]


class AddkGemmArgs(C.Structure):
    _fields_ = [("A", C.c_void_p), ("lda", C.c_int32), ("B", C.c_void_p), ("ldb", C.c_int32), ("C", C.c_void_p),
                ("ldc", C.c_int32), ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("bias", C.c_void_p),
                ("a_mean", C.c_void_p), ("a_std", C.c_void_p), ("relu_mask_src", C.c_void_p), ("ld_mask", C.c_int32),
                ("trans_a", C.c_int32), ("trans_b", C.c_int32), ("relu", C.c_int32), ("split_k", C.c_int32),
                ("accumulate", C.c_int32), ("slab_stride", C.c_int64),
                ("A16", C.c_void_p), ("B16", C.c_void_p), ("C16", C.c_void_p),
                ("a16_plane", C.c_int64), ("b16_plane", C.c_int64), ("a_amax", C.c_void_p), ("b_amax", C.c_void_p),
                ("a16_ready", C.c_int32), ("b16_ready", C.c_int32), ("c_amax", C.c_void_p), ("c16_plane", C.c_int64),
                ("relu_mask_src16", C.c_void_p), ("no_f32", C.c_int32),
                ("relu_bits_out", C.c_void_p), ("relu_bits_in", C.c_void_p), ("ld_bits", C.c_int32),
                ("colsum_partials", C.c_void_p)]


class AddkError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libaddk.so once.  Raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AddkError("libaddk.so not found at %s -- build it with add_gym_b200/csrc/Makefile; "
                            "there is no CPU fallback for the hot path" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.addk_last_error.restype = C.c_char_p
        _lib.addk_launch_count.restype = C.c_longlong
        _lib.addk_launch_count.argtypes = [C.c_int]
    return _lib


def has_legacy_kernels():
    """True when the loaded library was built with the superseded tf32 / tf32x3 kernels (make LEGACY=1)."""
    return bool(lib().addk_build_flags() & 1)


def check(rc, what):
    if rc != 0:
        raise AddkError("%s failed (code %d): %s" % (what, rc, lib().addk_last_error().decode()))


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  Refuses host tensors: no CPU path exists."""
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise AddkError("add_gym_b200 kernels need CUDA tensors (got a %s tensor)" % t.device)
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count(reset=False):
    return int(lib().addk_launch_count(1 if reset else 0))


def ctx_field_names():
    """(ptr names, int names, f64 names) in declaration order, parsed from csrc/ctx_fields.h."""
    ptrs, ints, f64s = [], [], []
    with open(CTX_FIELDS_PATH, "r") as f:
        for line in f:
            line = line.split("//")[0]
            for kind, name in re.findall(r"ADDK_(PTR|INT|F64)\((\w+)\)", line):
                if name == "n":
                    continue
                {"PTR": ptrs, "INT": ints, "F64": f64s}[kind].append(name)
    return ptrs, ints, f64s


class UpdateCtx:
    """Host-side addk_update_ctx: a plain struct of device pointers and scalars (csrc/mlp.cu)."""

    def __init__(self, ptr_values, int_values, f64_values):
        L = lib()
        pn, inn, fn = ctx_field_names()
        missing = [n for n in pn if n not in ptr_values] + [n for n in inn if n not in int_values] + \
                  [n for n in fn if n not in f64_values]
        if missing:
            raise AddkError("update ctx: missing fields %s" % missing)
        self._keep = [ptr_values[n] for n in pn]
        pa = (C.c_void_p * len(pn))(*[ptr(ptr_values[n]).value or 0 for n in pn])
        ia = (C.c_int64 * len(inn))(*[int(int_values[n]) for n in inn])
        fa = (C.c_double * len(fn))(*[float(f64_values[n]) for n in fn])
        self.size = L.addk_update_ctx_size()
        self.buf = C.create_string_buffer(self.size)
        check(L.addk_update_ctx_init(self.buf, pa, len(pn), ia, len(inn), fa, len(fn)), "addk_update_ctx_init")
        self.ints = dict(int_values)
        self.f64s = dict(f64_values)
        self.ptrs = dict(ptr_values)

    def rebuild(self, **changes):
        iv, fv, pv = dict(self.ints), dict(self.f64s), dict(self.ptrs)
        for k, v in changes.items():
            (iv if k in iv else fv if k in fv else pv)[k] = v
        return UpdateCtx(pv, iv, fv)


# ---- exchange step over NVLink peer memory (csrc/p2p.cu) -------------------------------------------------------------
class _ForeignCuda:
    """__cuda_array_interface__ holder: lets torch view device memory this library allocated (no copy, no ownership)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def foreign_tensor(ptr, n, dtype, device):
    typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
    return torch.as_tensor(_ForeignCuda(ptr, (n,), typestr), device=device)


class P2PExchange:
    """Peer-mapped flat gradient / parameter vectors of all ranks of one box + the flags of the fused exchange kernel.

    Every rank cudaMallocs its three buffers, the 64-byte cudaIpc handles travel once through torch.distributed, and each
    rank maps its peers' buffers.  `grad` / `param` are torch views of THIS rank's buffers (the model's flat vectors are
    rebound to them); `step()` launches addk_p2p_adamw."""

    FLAG_WORDS = 16

    def __init__(self, num_params, device, dist):
        L = lib()
        self.rank, self.world, self.n = dist.get_rank(), dist.get_world_size(), int(num_params)
        if self.world > 8:
            raise AddkError("the peer-memory exchange kernel addresses at most 8 ranks")
        own = []
        for nbytes in (4 * self.n, 4 * self.n, 4 * self.FLAG_WORDS):
            p = C.c_void_p()
            check(L.addk_p2p_alloc(C.c_longlong(nbytes), C.byref(p)), "addk_p2p_alloc")
            own.append(p.value)
        handles = []
        for p in own:
            h = (C.c_ubyte * 64)()
            check(L.addk_p2p_export(C.c_void_p(p), h), "addk_p2p_export")
            handles.append(bytes(h))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, handles)
        ptrs = [[0] * self.world for _ in range(3)]
        for r, hs in enumerate(everyone):
            for k in range(3):
                if r == self.rank:
                    ptrs[k][r] = own[k]
                else:
                    q = C.c_void_p()
                    hb = (C.c_ubyte * 64).from_buffer_copy(hs[k])
                    check(L.addk_p2p_open(hb, C.byref(q)), "addk_p2p_open (rank %d)" % r)
                    ptrs[k][r] = q.value
        self._own = own
        self.grad_ptrs = (C.c_void_p * self.world)(*ptrs[0])
        self.param_ptrs = (C.c_void_p * self.world)(*ptrs[1])
        self.flag_ptrs = (C.c_void_p * self.world)(*ptrs[2])
        self.grad = foreign_tensor(own[0], self.n, torch.float32, device)
        self.param = foreign_tensor(own[1], self.n, torch.float32, device)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=device)
        dist.barrier()      # every mapping exists before anyone launches

    def shard(self):
        """[begin, end) of the parameter shard whose AdamW moments this rank keeps current."""
        n4 = (self.n + 3) // 4
        per = (n4 + self.world - 1) // self.world
        return min(self.n, 4 * self.rank * per), min(self.n, 4 * (self.rank + 1) * per)

    def step(self, exp_avg, exp_avg_sq, step, lr, betas, eps, weight_decay):
        rc = lib().addk_p2p_adamw(stream(), C.c_int(self.rank), C.c_int(self.world), self.grad_ptrs, self.param_ptrs, self.flag_ptrs,
                                  ptr(exp_avg), ptr(exp_avg_sq), C.c_longlong(self.n), C.c_int(step), C.c_double(lr),
                                  C.c_double(betas[0]), C.c_double(betas[1]), C.c_double(eps), C.c_double(weight_decay),
                                  ptr(self.ticket), C.c_int(0))
        check(rc, "addk_p2p_adamw")
