"""Rollout storage of the B200 path: ``[T, N, ...]`` device tensors that the fused kernels fill row by row, plus the
minibatch index stream of the update.

The public method names are those of the reference container (add_gym/learning/experience_buffer.py:4-113) because
``ADDAgent`` subclasses and reference code call them; what sits behind them is different:

* the kernels write row ``t`` of seven keys themselves (csrc/step.cu, csrc/mlp.cu) -- ``record`` only serves keys a
  subclass adds;
* the update never gathers per key in Python: it asks for an index window (``sample_indices``) and the gather of the
  eight keys it needs happens inside ``addk_update_minibatch``;
* the index window is ONE native launch (``addk_perm_window``: ``perm[(head + i) mod L] mod sample_count``) instead of
  the reference's slice / cat / remainder chain.  The draw order is the reference's -- a device permutation of T*N
  entries, drawn twice at construction, consumed n entries at a time and re-drawn IN PLACE when a request wraps, so the
  tail of a wrapping request already shows the re-drawn values (the reference's tail slice is a view) -- which is what
  lets the parity tests replay the oracle's permutations draw for draw.
"""
import ctypes as C
from collections import namedtuple

import torch

from . import _lib

_Key = namedtuple("_Key", "grid flat")       # the [T, N, ...] tensor and its [T*N, ...] view


class ExperienceBuffer:
    def __init__(self, buffer_length, batch_size, device, randperm_fn=None):
        self._T, self._N = int(buffer_length), int(batch_size)
        self._device = device
        self._keys = {}
        self._row = 0                 # next row to be written (the reference's buffer head)
        self._seen = 0                # samples recorded since the last clear()
        draw = randperm_fn or (lambda n: torch.randperm(n, device=device, dtype=torch.long))
        self._draw = lambda: draw(self._T * self._N)
        self._perm = self._draw()     # (first of the two draws the reference makes at construction)
        self._cursor = 0
        self._redraw()

    # ---- storage ------------------------------------------------------------------------------------------------
    def add_buffer(self, name, buffer):
        if name in self._keys:
            raise KeyError("experience key %r exists" % name)
        if buffer.dim() < 2 or tuple(buffer.shape[:2]) != (self._T, self._N):
            raise ValueError("experience key %r must be [%d, %d, ...], got %s" % (name, self._T, self._N, tuple(buffer.shape)))
        self._keys[name] = _Key(buffer, buffer.view(self._T * self._N, *buffer.shape[2:]))

    def get_data(self, name):
        return self._keys[name].grid

    def get_data_flat(self, name):
        return self._keys[name].flat

    def set_data(self, name, data):
        grid = self._keys[name].grid
        if tuple(data.shape[:2]) != tuple(grid.shape[:2]):
            raise ValueError("set_data(%r): leading shape %s != %s" % (name, tuple(data.shape[:2]), tuple(grid.shape[:2])))
        grid.copy_(data)

    def set_data_flat(self, name, data):
        flat = self._keys[name].flat
        if data.shape[0] != flat.shape[0]:
            raise ValueError("set_data_flat(%r): %d rows != %d" % (name, data.shape[0], flat.shape[0]))
        flat.copy_(data)

    def record(self, name, data):
        """Row `head` of a key the kernels do not write themselves (subclass keys)."""
        if data.shape[0] != self._N:
            raise ValueError("record(%r): %d envs != %d" % (name, data.shape[0], self._N))
        self._keys[name].grid[self._row].copy_(data)

    # ---- cursor over the rows -----------------------------------------------------------------------------------
    def inc(self):
        self._row = (self._row + 1) % self._T
        self._seen += self._N

    def reset(self):
        self._row = 0
        self._redraw()

    def clear(self):
        self.reset()
        self._seen = 0

    def get_buffer_head(self):
        return self._row

    def get_total_samples(self):
        return self._seen

    def get_sample_count(self):
        return min(self._seen, self._T * self._N)

    # ---- minibatch index stream ---------------------------------------------------------------------------------
    def _redraw(self):
        self._perm.copy_(self._draw())         # in place: outstanding views follow, as in the reference
        self._cursor = 0

    def sample_indices(self, n):
        """The next n minibatch indices as a fresh contiguous int64 tensor (one launch)."""
        length = self._perm.shape[0]
        if not 0 < n <= length:
            raise ValueError("sample_indices(%d): the permutation has %d entries" % (n, length))
        start = self._cursor
        if start + n > length:                 # the request wraps: re-draw first, the window then reads the new values
            self._redraw()
            self._cursor = start + n - length
        else:
            self._cursor = start + n
        out = torch.empty(n, dtype=torch.long, device=self._perm.device)
        _lib.check(_lib.lib().addk_perm_window(_lib.stream(), _lib.ptr(self._perm), C.c_longlong(length), C.c_longlong(start),
                                               C.c_int(n), C.c_longlong(self.get_sample_count()), _lib.ptr(out)),
                   "addk_perm_window")
        return out

    def sample(self, n):
        """Per-key gather of one minibatch (API compatibility; the update gathers natively)."""
        idx = self.sample_indices(n)
        return {name: key.flat.index_select(0, idx) for name, key in self._keys.items()}
