"""Running statistics, drop-in for the reference's ``Normalizer`` (normalizer.py:7-162) and
``DiffNormalizer`` (diff_normalizer.py:6-86): same buffers / state-dict keys (``_count`` int64[1],
``_mean``, ``_std`` | ``_mean_abs``), same update rule.

``record`` accumulates column sums in fp64 on the device (addk_column_stats); the agent's fast path
records the whole [T*N, dim] experience buffer once per iteration instead of once per step, which is the
same sum.  ``update`` merges them into the running moments with one small kernel; under
torch.distributed the fp64 sums and the count are all-reduced first in ONE message (the reference sends
three, normalizer.py:41-58).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


class Normalizer(torch.nn.Module):
    def __init__(self, shape, device, init_mean=None, init_std=None, min_std=1e-4, clip=np.inf, dtype=torch.float):
        super().__init__()
        self._min_var = min_std * min_std
        self._clip = clip
        self.dtype = dtype
        shape = tuple(shape)
        assert len(shape) == 1, "the B200 path normalizes flat feature vectors"
        self._dim = shape[0]
        P = lambda t: torch.nn.Parameter(t, requires_grad=False)
        self._count = P(torch.zeros([1], device=device, dtype=torch.long))
        self._mean = P(torch.zeros(shape, device=device, dtype=dtype))
        self._std = P(torch.ones(shape, device=device, dtype=dtype))
        if init_mean is not None:
            self._mean[:] = init_mean
        if init_std is not None:
            self._std[:] = init_std
        self._mean_sq = None
        self._new_count = 0
        self._sums = torch.zeros(2 * self._dim + 1, dtype=torch.float64, device=device)

    def record(self, x):
        x = x.reshape(-1, self._dim)
        if not x.is_contiguous():
            x = x.contiguous()
        self._new_count += x.shape[0]
        rc = _lib.lib().addk_column_stats(_lib.stream(), _lib.ptr(x), None, C.c_longlong(x.shape[0]), C.c_int(self._dim),
                                          C.c_int(0), _lib.ptr(self._sums))
        _lib.check(rc, "addk_column_stats")

    def update(self):
        if self._mean_sq is None:
            self._mean_sq = (torch.square(self._std) + torch.square(self._mean)).type(self.dtype)
        new_count = self._new_count
        if dist.is_available() and dist.is_initialized():
            self._sums[-1] = float(new_count)
            dist.all_reduce(self._sums, op=dist.ReduceOp.SUM)
            new_count = int(round(self._sums[-1].item()))
        if new_count == 0:
            return
        rc = _lib.lib().addk_normalizer_update(
            _lib.stream(), _lib.ptr(self._sums), C.c_double(float(new_count)), C.c_int(self._dim), _lib.ptr(self._count),
            _lib.ptr(self._mean), _lib.ptr(self._mean_sq), _lib.ptr(self._std), C.c_float(self._min_var))
        _lib.check(rc, "addk_normalizer_update")
        self._new_count = 0
        self._sums.zero_()

    def get_shape(self):
        return self._mean.shape

    def get_count(self):
        return self._count

    def get_mean(self):
        return self._mean

    def get_std(self):
        return self._std

    def set_mean_std(self, mean, std):
        self._mean[:] = mean
        self._std[:] = std
        self._mean_sq = (torch.square(self._std) + torch.square(self._mean)).type(self.dtype)

    def normalize(self, x):
        norm_x = (x - self._mean) / self._std
        if np.isfinite(self._clip):
            norm_x = torch.clamp(norm_x, -self._clip, self._clip)
        return norm_x.type(self.dtype)

    def unnormalize(self, norm_x):
        return (norm_x * self._std + self._mean).type(self.dtype)


class DiffNormalizer(torch.nn.Module):
    def __init__(self, shape, device, init_mean=None, min_diff=1e-4, clip=np.inf, dtype=torch.float):
        super().__init__()
        self._min_diff = min_diff
        self._clip = clip
        self.dtype = dtype
        shape = tuple(shape)
        self._dim = shape[0]
        P = lambda t: torch.nn.Parameter(t, requires_grad=False)
        self._count = P(torch.zeros([1], device=device, dtype=torch.long))
        self._mean_abs = P(torch.ones(shape, device=device, dtype=dtype))
        if init_mean is not None:
            self._mean_abs[:] = init_mean
        self._new_count = 0
        self._sums = torch.zeros(self._dim, dtype=torch.float64, device=device)

    def record(self, x):
        """x = demo - agent (reference signature).  Prefer record_pair to skip the temporary."""
        zero = torch.zeros_like(x)
        self.record_pair(x, zero)

    def record_pair(self, demo, agent):
        a = demo.reshape(-1, self._dim).contiguous()
        b = agent.reshape(-1, self._dim).contiguous()
        self._new_count += a.shape[0]
        rc = _lib.lib().addk_column_stats(_lib.stream(), _lib.ptr(a), _lib.ptr(b), C.c_longlong(a.shape[0]),
                                          C.c_int(self._dim), C.c_int(1), _lib.ptr(self._sums))
        _lib.check(rc, "addk_column_stats")

    def update(self):
        # per-rank, no all-reduce and no zero guard, as in the reference (diff_normalizer.py:33-45)
        rc = _lib.lib().addk_diff_normalizer_update(
            _lib.stream(), _lib.ptr(self._sums), C.c_double(float(self._new_count)), C.c_int(self._dim),
            _lib.ptr(self._count), _lib.ptr(self._mean_abs))
        _lib.check(rc, "addk_diff_normalizer_update")
        self._new_count = 0
        self._sums.zero_()

    def get_shape(self):
        return self._mean_abs.shape

    def get_count(self):
        return self._count

    def get_abs_mean(self):
        return self._mean_abs

    def normalize(self, x):
        norm_x = x / torch.clamp_min(self._mean_abs, self._min_diff)
        if np.isfinite(self._clip):
            norm_x = torch.clamp(norm_x, -self._clip, self._clip)
        return norm_x.type(self.dtype)

    def unnormalize(self, norm_x):
        return (norm_x * torch.clamp_min(self._mean_abs, self._min_diff)).type(self.dtype)
